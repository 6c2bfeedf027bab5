"""Inference runtime: the call a user makes.

``Detector``         one GPU: uint8 RGB + IR images (host or device) -> padded detections.
                     Mirrors the reference's evaluation loop (basics/test.py:124-152): /255,
                     ``model(img, ir, input_mode)``, ``non_max_suppression``.
``ShardedDetector``  one process per GPU: images are sharded by index across ranks, weights are
                     replicated, the forward has no collective; the only exchange is one
                     all-gather of the fixed-shape padded detections ([B_local, 300, 6] fp32 +
                     [B_local] int32, 7.2 KB per image).  The NMS kernel writes straight into the
                     all-gather send buffer, so there is no pack step.
"""
import os

import torch
import torch.distributed as dist

from . import ops
from .basics.models.model import Model

MAX_DET = 300
DEFAULT_CFG = os.path.join(os.path.dirname(os.path.abspath(__file__)), "models", "model.yaml")


def shard_range(n_items, rank, world):
    """Contiguous split of ``n_items`` image indices over ``world`` ranks (first ranks get the remainder)."""
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


class DetectionBuffer:
    """Flat fp32 communication buffer: [B*max_det*6 detection floats | B int32 counts (bit-cast)].
    ``det`` and ``counts`` are views into it; NMS writes them in place."""

    def __init__(self, batch, device, max_det=MAX_DET):
        self.batch, self.max_det = batch, max_det
        self.flat = torch.zeros(batch * max_det * 6 + batch, dtype=torch.float32, device=device)
        self.det = self.flat[: batch * max_det * 6].view(batch, max_det, 6)
        self.counts = self.flat[batch * max_det * 6:].view(torch.int32)

    @staticmethod
    def split(flat, batch, max_det=MAX_DET):
        n = batch * max_det * 6
        return flat[:n].view(batch, max_det, 6), flat[n:n + batch].view(torch.int32)


class GatheredDetections:
    """The all-gathered flat buffers of every rank, split lazily: ``det`` [world, B, max_det, 6] and ``counts`` [world, B] are
    strided VIEWS of the receive buffer (no re-assembly kernels); ``tensors()`` makes the contiguous [world*B, ...] copies."""

    def __init__(self, flat, world, batch, max_det=MAX_DET, event=None):
        self.flat, self.world, self.batch, self.max_det, self.event = flat, world, batch, max_det, event
        rows = flat.view(world, -1)
        n = batch * max_det * 6
        self.det = rows[:, :n].unflatten(1, (batch, max_det, 6))
        self.counts = rows[:, n:n + batch].view(torch.int32)

    def wait(self, stream=None):
        """Makes ``stream`` (default: the current one) wait for the gather; call before reading the views on the device."""
        if self.event is not None:
            (stream or torch.cuda.current_stream(self.flat.device)).wait_event(self.event)
        return self

    def tensors(self):
        self.wait()
        return self.det.reshape(-1, self.max_det, 6), self.counts.reshape(-1)


def allgather_detections(buf, group=None):
    """All-gathers every rank's DetectionBuffer.  Returns (det [world*B,300,6], counts [world*B]) in rank order == global
    image order for contiguous sharding.  The ranks must hold EQUAL local batches (all_gather_into_tensor has one block size):
    an uneven split from shard_range must be padded by the caller; checked here."""
    world = dist.get_world_size(group)
    sizes = torch.tensor([buf.batch], dtype=torch.int64, device=buf.flat.device)
    if world > 1:
        lo, hi = sizes.clone(), sizes.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=group)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=group)
        if int(lo) != int(hi):
            raise ValueError(f"allgather_detections needs equal local batches on every rank (got {int(lo)}..{int(hi)}): pad the shards")
    recv = torch.empty(world * buf.flat.numel(), dtype=buf.flat.dtype, device=buf.flat.device)
    dist.all_gather_into_tensor(recv, buf.flat, group=group)
    return GatheredDetections(recv, world, buf.batch, buf.max_det).tensors()


class Detector:
    """RGB+IR detector on one GPU.  ``dtype`` is the storage / tensor-core operand type of the
    backbone and head (bf16 by default); softmax, LayerNorm statistics, Detect decode and NMS are fp32."""

    def __init__(self, cfg=DEFAULT_CFG, state_dict=None, device="cuda", dtype=torch.bfloat16, conf_thres=0.25,
                 iou_thres=0.45, multi_label=False, agnostic=False, classes=None, nc=8, seed=0, cuda_graph=True):
        self.device = torch.device(device)
        self.dtype = dtype
        if classes is not None and self.device.type == "cuda":
            # uploaded once: ops.nms then enqueues no host-to-device copy (none is allowed inside CUDA-graph capture)
            classes = ops.class_filter(classes, torch.device("cuda", torch.cuda.current_device())
                                       if self.device.index is None else self.device)
        self.nms_args = dict(conf_thres=conf_thres, iou_thres=iou_thres, multi_label=multi_label, agnostic=agnostic,
                             classes=classes)
        torch.manual_seed(seed)
        model = Model(cfg, input_mode="RGB+IR", ch_steam=3, ch=128, nc=nc)
        if state_dict is not None:
            self._load_checked(model, state_dict)
        # like the reference's inference loader (models/experimental.py:118-120): fold BatchNorm into the convs
        self.model = model.eval().fuse().to(self.device, dtype)
        for m in self.model.modules():
            if hasattr(m, "want_raw"):                  # the runtime reads only the decoded predictions
                m.want_raw = False
            if hasattr(m, "anchor_grid"):               # Detect decodes in fp32: its anchors must not be rounded to bf16
                m.anchors = m.anchors.float()
                m.anchor_grid = m.anchor_grid.float()
        self.copy_stream = torch.cuda.Stream(self.device) if self.device.type == "cuda" else None
        self._bufs = {}
        # one CUDA graph per input shape: the ~1400 launches of a step are replayed with one driver call
        self.cuda_graph = bool(cuda_graph) and self.device.type == "cuda"
        self._graphs = {}

    # buffers whose shape follows the token grid, not the weights: the only keys a checkpoint may lack / add
    _GRID_KEYS = ("attn_mask", "relative_position_index", "pos_embed")

    @classmethod
    def _load_checked(cls, model, state_dict):
        """load_state_dict that fails on any missing / unexpected key except the resolution-dependent buffers: a
        mis-keyed checkpoint must not leave random-init weights behind silently."""
        res = model.load_state_dict(state_dict, strict=False)
        bad = [k for k in list(res.missing_keys) + list(res.unexpected_keys) if not k.endswith(cls._GRID_KEYS)]
        if bad:
            raise RuntimeError(f"checkpoint does not match the model: {len(bad)} missing / unexpected keys, e.g. {bad[:5]}")
        return res

    def load_state_dict(self, state_dict):
        """Replaces the weights of the (fused, device-resident) model.  Captured CUDA graphs hold pointers to cached folded
        weights, so they are dropped and re-captured on the next call."""
        res = self._load_checked(self.model, state_dict)
        self._graphs.clear()
        return res

    def buffer(self, batch):
        if batch not in self._bufs:
            self._bufs[batch] = DetectionBuffer(batch, self.device)
        return self._bufs[batch]

    @torch.no_grad()
    def predict(self, rgb_u8, ir_u8):
        """Device uint8 [B,3,H,W] x2 -> decoded predictions [B, R, 5+nc] fp32 (the reference's ``out``)."""
        if rgb_u8.dtype == torch.uint8 and ir_u8.dtype == torch.uint8:
            pred, _, _ = self.model(rgb_u8, ir_u8, "RGB+IR")      # /255 and the IR channel-0 pick happen in the front-end kernel
            return pred
        x = rgb_u8.to(self.dtype).div_(255.0)
        ir = ir_u8[:, 0:1].to(self.dtype).div_(255.0)     # the detector reads IR channel 0 only (model.py:192)
        pred, _, _ = self.model(x, ir, "RGB+IR")
        return pred

    @torch.no_grad()
    def _detect_eager(self, rgb_u8, ir_u8, buf):
        pred = self.predict(rgb_u8, ir_u8)
        ops.nms(pred, out=buf.det, counts=buf.counts, **self.nms_args)
        return buf

    def _graph_for(self, rgb_u8, ir_u8):
        """Captures (once per input shape) the whole step on static input / output buffers."""
        key = (tuple(rgb_u8.shape), tuple(ir_u8.shape))
        entry = self._graphs.get(key)
        if entry is None:
            srgb, sir = rgb_u8.clone(), ir_u8.clone()
            buf = DetectionBuffer(rgb_u8.shape[0], self.device)
            cur = torch.cuda.current_stream(self.device)
            side = torch.cuda.Stream(self.device)
            side.wait_stream(cur)
            with torch.cuda.stream(side):          # lazy initialisation (allocator, weight caches, function attributes)
                for _ in range(2):
                    self._detect_eager(srgb, sir, buf)
            cur.wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            n0 = ops.launch_count()
            with torch.cuda.graph(graph):
                self._detect_eager(srgb, sir, buf)
            entry = (graph, srgb, sir, buf, ops.launch_count() - n0)
            self._graphs[key] = entry
        return entry

    def launches_per_step(self, rgb_u8, ir_u8):
        """sodt kernels one step launches (counted while the step's CUDA graph was captured; 0 if graphs are off)."""
        return self._graph_for(rgb_u8, ir_u8)[4] if self.cuda_graph else 0

    @torch.no_grad()
    def detect_device(self, rgb_u8, ir_u8, buf=None):
        """Device uint8 inputs -> DetectionBuffer (device resident, no host sync).  Without ``buf`` the returned buffer
        is reused by the next call with the same batch size."""
        if (self.cuda_graph and rgb_u8.dtype == torch.uint8 and ir_u8.dtype == torch.uint8 and not ops.kernel_timing_enabled()
                and not torch.cuda.is_current_stream_capturing()):
            try:
                graph, srgb, sir, gbuf, _ = self._graph_for(rgb_u8, ir_u8)
            except RuntimeError as e:       # capture not possible here (still the CUDA path: every kernel is launched eagerly)
                import warnings
                warnings.warn(f"CUDA graph capture failed ({e}); launching the step eagerly")
                self.cuda_graph = False
                torch.cuda.synchronize(self.device)
                return self._detect_eager(rgb_u8, ir_u8, buf or self.buffer(rgb_u8.shape[0]))
            srgb.copy_(rgb_u8)
            sir.copy_(ir_u8)
            graph.replay()
            if buf is None:
                return gbuf
            buf.flat.copy_(gbuf.flat)
            return buf
        return self._detect_eager(rgb_u8, ir_u8, buf or self.buffer(rgb_u8.shape[0]))

    def detect(self, rgb_u8_host, ir_u8_host):
        """Host uint8 images (pinned for async copies) -> (det [B,300,6], counts [B]) on the host."""
        rgb = rgb_u8_host.to(self.device, non_blocking=True)
        ir = ir_u8_host.to(self.device, non_blocking=True)
        buf = self.detect_device(rgb, ir)
        flat = buf.flat.to("cpu")
        return DetectionBuffer.split(flat, buf.batch, buf.max_det)

    def _rings(self, batch, depth=3, world=1):
        key = ("ring", batch, world)
        if key not in self._bufs:
            dev = [DetectionBuffer(batch, self.device) for _ in range(depth)]
            n = dev[0].flat.numel() * world
            gathered = [torch.empty(n, dtype=torch.float32, device=self.device) for _ in range(depth)] if world > 1 else None
            host = [torch.empty(n, dtype=torch.float32).pin_memory() for _ in range(depth)]
            self._bufs[key] = (dev, host, gathered)
        return self._bufs[key]

    def detect_stream(self, batches, group=None, world=1):
        """Pipelined inference over an iterable of (rgb_u8_host, ir_u8_host) pinned batches: the host->device copy of
        batch i+1 runs on a copy stream while batch i is computed, and the detections of batch i are read back while
        batch i+1 is computed.  Device / pinned host result buffers come from small preallocated rings (a yielded result
        stays valid until two more results have been taken).  Yields (det [B,300,6], counts [B]) host tensors in order.
        With ``world`` > 1 (ShardedDetector.detect_stream) every batch is this rank's shard: the ranks' detections are
        all-gathered on the device before the read-back and the yielded tensors cover all ``world * B`` images."""
        compute = torch.cuda.current_stream(self.device)
        copy = self.copy_stream

        def upload(pair):
            with torch.cuda.stream(copy):
                rgb = pair[0].to(self.device, non_blocking=True)
                ir = pair[1].to(self.device, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy)
            return rgb, ir, ev

        def split(flat, batch):
            if world == 1:
                return DetectionBuffer.split(flat, batch, MAX_DET)
            n = flat.numel() // world
            parts = [DetectionBuffer.split(flat[r * n:(r + 1) * n], batch, MAX_DET) for r in range(world)]
            return torch.cat([d for d, _ in parts]), torch.cat([c for _, c in parts])

        it = iter(batches)
        nxt = next(it, None)
        staged = upload(nxt) if nxt is not None else None
        pending = None                                   # (host flat buffer, event, batch) of the previous batch
        i = 0
        while staged is not None:
            rgb, ir, ev = staged
            nxt = next(it, None)
            staged = upload(nxt) if nxt is not None else None      # overlaps with the compute below
            dev_ring, host_ring, gather_ring = self._rings(rgb.shape[0], world=world)
            host = host_ring[i % len(host_ring)]
            compute.wait_event(ev)
            if pending is not None:
                compute.wait_event(pending[1])         # the previous read-back / gather has finished with the step's output buffer
            # graph path: the replayed NMS writes into the graph's own buffer, which is read back / gathered in place (no pack copy)
            buf = self.detect_device(rgb, ir)
            rgb.record_stream(compute)
            ir.record_stream(compute)
            src = buf.flat
            if world > 1:                                           # the only collective of the path: padded detections
                src = gather_ring[i % len(gather_ring)]
                dist.all_gather_into_tensor(src, buf.flat, group=group)
            done = torch.cuda.Event()
            done.record(compute)
            with torch.cuda.stream(copy):
                copy.wait_event(done)
                host.copy_(src, non_blocking=True)
                hev = torch.cuda.Event()
                hev.record(copy)
            if pending is not None:
                pending[1].synchronize()
                yield split(pending[0], pending[2])
            pending = (host, hev, buf.batch)
            i += 1
        if pending is not None:
            pending[1].synchronize()
            yield split(pending[0], pending[2])

    def __call__(self, rgb_u8_host, ir_u8_host):
        det, counts = self.detect(rgb_u8_host, ir_u8_host)
        return [det[i, :n] for i, n in enumerate(counts.tolist())]


class ShardedDetector:
    """Data-parallel inference over the ranks of an initialised process group (one process per GPU)."""

    def __init__(self, detector, group=None):
        self.detector = detector
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.comm_stream = torch.cuda.Stream(detector.device) if detector.device.type == "cuda" else None
        self._recv = {}                 # ring of receive buffers per batch size
        self._step = 0
        self._pending = None            # gather of the previous step (its send buffer is the next step's NMS output)

    def local_slice(self, n_images):
        return shard_range(n_images, self.rank, self.world)

    @torch.no_grad()
    def detect_device(self, rgb_u8_local, ir_u8_local):
        """Each rank passes ITS shard (equal sizes); returns the gathered detections of all ranks as a GatheredDetections
        (views of the receive buffer; ``.tensors()`` for contiguous copies).  The NMS kernel of the (graph-replayed) step writes
        straight into the buffer NCCL sends; the all-gather runs on a side stream and overlaps the next step, which only waits
        for it before its own NMS output could overwrite the send buffer."""
        det = self.detector
        if self.comm_stream is None:                          # host tensors (gloo; the CPU tests of the host logic): a synchronous gather
            buf = det.detect_device(rgb_u8_local, ir_u8_local)
            recv = torch.empty(self.world * buf.flat.numel(), dtype=torch.float32, device=buf.flat.device)
            dist.all_gather_into_tensor(recv, buf.flat, group=self.group)
            self._step += 1
            return GatheredDetections(recv, self.world, buf.batch, buf.max_det)
        compute = torch.cuda.current_stream(det.device)
        if self._pending is not None:                     # the previous gather still reads the send buffer this step rewrites
            compute.wait_event(self._pending)
        buf = det.detect_device(rgb_u8_local, ir_u8_local)
        n = buf.flat.numel()
        ring = self._recv.setdefault(buf.batch, [torch.empty(self.world * n, dtype=torch.float32, device=det.device) for _ in range(3)])
        recv = ring[self._step % len(ring)]
        self._step += 1
        ready = torch.cuda.Event()
        ready.record(compute)
        with torch.cuda.stream(self.comm_stream):
            self.comm_stream.wait_event(ready)
            dist.all_gather_into_tensor(recv, buf.flat, group=self.group)
            done = torch.cuda.Event()
            done.record(self.comm_stream)
        recv.record_stream(self.comm_stream)
        self._pending = done
        return GatheredDetections(recv, self.world, buf.batch, buf.max_det, event=done)

    def detect_stream(self, local_batches):
        """Pipelined host API (see Detector.detect_stream): each rank feeds its own shard of every batch and receives the
        detections of all ranks' images, in rank order."""
        return self.detector.detect_stream(local_batches, group=self.group, world=self.world)
