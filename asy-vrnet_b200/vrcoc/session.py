"""Serving-side product API around the model: the step right before and right after the hot path.

`InferenceSession` owns what `yolo.py:123-146` / `deeplab.py:149-167` do per frame around `net(images, radar)` on the device
side — host->device copy of the pre-processed frame batch, the forward, box decoding, the per-pixel class map, device->host
copy of the results — as a fixed pipeline: the whole forward (incl. decode and arg-max) is captured once in a CUDA graph per
pipeline slot, inputs arrive from pinned host memory on a copy stream, results leave on another one, and with two slots the
copies of batch i+1 / i-1 run under the forward of batch i.  With `concurrent=True` (default) every slot also replays its graph on
its OWN compute stream, so the forwards of consecutive batches overlap as well: at batch 8 a forward is a chain of ~230 launches
of which 40 % run alone on at most half of the 148 SMs (profiles/r02_timeline.txt); the second batch in flight fills them.  Each
batch is still computed on its own (batch-global statistics such as data_normal's min / max are per batch); per-batch latency
grows, frames per second go up.

`decode_outputs` is the reference's utils/utils_bbox.py:32-84 as one kernel (`vrcoc_decode_outputs`).
"""
import os

import torch

from . import ops
from ._lib import VrcocError, check, lib


def decode_outputs(outputs, input_shape):
    """[B, 5+nc, h, w] x 3 detection maps -> [B, sum hw, 5+nc] fp32 (normalised cx, cy, w, h, objectness, class scores);
    reference utils/utils_bbox.py:32-84, one launch."""
    p3, p4, p5 = (o.contiguous() for o in outputs)
    if not p3.is_cuda:
        raise VrcocError("vrcoc decode_outputs needs CUDA tensors (no CPU fallback exists)")
    B, CH = p3.shape[:2]
    N = sum(o.shape[2] * o.shape[3] for o in (p3, p4, p5))
    out = torch.empty(B, N, CH, device=p3.device, dtype=torch.float32)
    check(lib.vrcoc_decode_outputs(p3.data_ptr(), p4.data_ptr(), p5.data_ptr(), ops._dt(p3), B, CH, p3.shape[2], p3.shape[3], p4.shape[2],
                                   p4.shape[3], p5.shape[2], p5.shape[3], int(input_shape[0]), int(input_shape[1]), out.data_ptr(),
                                   ops._stream()), "decode_outputs")
    return out


class InferenceSession:
    """Fixed-shape inference pipeline for an `EfficientVRNet`-shaped model (forward(images, radars) -> (det maps, seg logits)).

        sess = InferenceSession(model, batch=8, img=512)
        boxes, classes = sess.run(images, radars)            # one batch, synchronous
        for images, radars in stream: sess.submit(images, radars); ... = sess.collect()     # pipelined, one batch of latency

    images [B,3,H,W], radars [B,4,H,W]: host tensors (pinned memory makes the copies asynchronous) or device tensors, any float
    dtype (converted to the model's).  Results per batch, as pinned host tensors owned by the session (valid until the slot is
    reused, i.e. for `slots` further submits): decoded detections [B, N, 5+nc] fp32 (or the three raw maps with decode=False) and
    the segmentation class map [B,H,W] uint8."""

    def __init__(self, model, batch, img=512, slots=2, decode=True, cuda_graph=True, device=None, concurrent=True):
        p = next(model.parameters())
        if not p.is_cuda and device is None:
            raise VrcocError("InferenceSession needs the model on a CUDA device (no CPU fallback exists)")
        self.device = torch.device(device) if device is not None else p.device
        self.dtype = p.dtype
        self.model = model.eval()
        self.batch, self.img, self.decode = batch, img, decode
        self.h2d, self.d2h = torch.cuda.Stream(self.device), torch.cuda.Stream(self.device)
        self.slots, self._next, self._pending = [], 0, []
        # concurrent: number of compute streams the slots' forwards rotate over (True = 2, measured best at batch 8: a third forward
        # in flight only thrashes L2; more SLOTS than streams still help the host side of the pipeline, which then runs further ahead)
        nstreams = 0 if (not concurrent or slots < 2) else min(slots, int(os.environ.get("VRCOC_SESSION_STREAMS", "2")) if concurrent is True else int(concurrent))
        self.concurrent = nstreams > 1
        self._cstreams = [torch.cuda.Stream(self.device) for _ in range(nstreams)] if self.concurrent else []
        self._nsub = 0
        with torch.cuda.device(self.device), torch.no_grad():
            for i in range(max(1, slots)):
                S = {"x": torch.zeros(batch, 3, img, img, device=self.device, dtype=self.dtype),
                     "r": torch.zeros(batch, 4, img, img, device=self.device, dtype=self.dtype), "graph": None,
                     # concurrent slots: own statistics arena (ops.sums_arena.lane), replayed on the session's compute streams
                     "lane": ops.sums_arena.new_lane() if self.concurrent else 0}
                lane, ops.sums_arena.lane = ops.sums_arena.lane, S["lane"]
                try:
                    for _ in range(2):                               # warm-up: memoised parameter views, lazy module state
                        outs = self._forward(S["x"], S["r"])
                    torch.cuda.synchronize(self.device)
                    if cuda_graph:
                        S["graph"] = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(S["graph"]):
                            outs = self._forward(S["x"], S["r"])
                finally:
                    ops.sums_arena.lane = lane
                S["outs"] = outs
                S["host"] = [torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in outs]
                S["in"], S["done"], S["out"] = torch.cuda.Event(), torch.cuda.Event(), torch.cuda.Event()
                S["done"].record(); S["out"].record()
                self.slots.append(S)
            torch.cuda.synchronize(self.device)

    def _forward(self, x, r):
        # the neck returns the class map itself (upsample + arg-max in one kernel) when it has the serving switch
        neck = getattr(self.model, "backbone", None)
        fused = neck is not None and hasattr(neck, "seg_class_map")
        if fused:
            neck.seg_class_map = True
        try:
            det, seg = self.model(x, r)
        finally:
            if fused:
                neck.seg_class_map = False
        cls = seg if seg.dtype == torch.uint8 else seg.argmax(dim=1).to(torch.uint8)
        if self.decode:
            return [decode_outputs(det, (self.img, self.img)), cls]
        return list(det) + [cls]

    def submit(self, images, radars, readback=True):
        """queue one batch; returns immediately (copies and forward are asynchronous).  readback=False leaves the results on the
        device (slot["outs"], valid until the slot is reused) and queues nothing for collect()."""
        S = self.slots[self._next]
        self._next = (self._next + 1) % len(self.slots)
        caller = torch.cuda.current_stream(self.device)
        main = self._cstreams[self._nsub % len(self._cstreams)] if self._cstreams else caller
        self._nsub += 1
        with torch.cuda.stream(self.h2d):
            if self._cstreams:
                self.h2d.wait_stream(caller)                    # device-resident inputs produced on the caller's stream
            self.h2d.wait_event(S["done"])                      # the slot's previous forward has consumed its inputs
            S["x"].copy_(images, non_blocking=True)
            S["r"].copy_(radars, non_blocking=True)
            S["in"].record(self.h2d)
        main.wait_event(S["in"])
        main.wait_event(S["out"])                               # the slot's previous results have left the device
        with torch.cuda.stream(main):
            if S["graph"] is not None:
                S["graph"].replay()
            else:
                lane, ops.sums_arena.lane = ops.sums_arena.lane, S["lane"]
                try:
                    with torch.no_grad():
                        for d, s in zip(S["outs"], self._forward(S["x"], S["r"])):
                            d.copy_(s)
                finally:
                    ops.sums_arena.lane = lane
            S["done"].record(main)
        if not readback:
            S["out"].record(main)
            return
        with torch.cuda.stream(self.d2h):
            self.d2h.wait_event(S["done"])
            for h, d in zip(S["host"], S["outs"]):
                h.copy_(d, non_blocking=True)
            S["out"].record(self.d2h)
        self._pending.append(S)

    def streams(self):
        """every stream the session queues work on (for callers that time it with events on their own stream)"""
        return [self.h2d, self.d2h] + list(self._cstreams)

    def collect(self):
        """results of the oldest submitted batch (blocks until its device->host copies are done)"""
        if not self._pending:
            raise VrcocError("InferenceSession.collect: nothing submitted")
        S = self._pending.pop(0)
        S["out"].synchronize()
        return S["host"]

    def run(self, images, radars):
        self.submit(images, radars)
        return self.collect()
