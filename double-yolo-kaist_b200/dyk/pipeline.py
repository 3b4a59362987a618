"""Software pipeline around the hot path for evaluate.py-style loops (reference evaluate.py:60-90: per batch
`imgs.to(device).float() / 255`, `model(v, l)[0]`, `non_max_suppression(...)`, results consumed on the host).

Three CUDA streams, everything still in program order per batch:
  copy stream : pinned host uint8 frames -> device (batch i+1 is uploaded while batch i computes)
  main stream : model forward (CUDA graph with the two modality lanes inside)
  nms stream  : batched NMS of batch i (one CTA per image, 16 of 148 SMs) overlaps the forward of batch i+1
Results come back one call later (`submit` returns the detections of the previous batch); `flush()` drains the last one.
With `to_host=True` the detections are also copied to pinned host memory on the nms stream and `submit` / `flush` return
those host tensors: reading results back with `.cpu()` on the main stream would make the host wait for the forward it has
just enqueued and serialise host and device once per batch (measured: 4 % of the end-to-end rate).
Nothing here computes: it only orders native launches with events.
"""
from __future__ import annotations

import torch

from build_utils.utils import nms_raw


class EvalPipeline:
    def __init__(self, model, conf_thres=0.01, iou_thres=0.6, multi_label=False, classes=None, agnostic=False, max_num=100,
                 to_host=False):
        self.model = model
        self.to_host = to_host
        self._host_ring, self._host_slot = [], 0     # pinned (out, counts) buffers; a result stays valid for two more submits
        self.args = (conf_thres, iou_thres, multi_label, classes, agnostic, max_num)
        dev = next(model.parameters()).device
        self.dev = dev
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.nms_stream = torch.cuda.Stream(device=dev)
        self.dual = "second_index" in model.net_info
        self._staged = None        # (v_dev, l_dev, event) uploaded ahead of time
        self._pending = None       # (out, counts, event) of the previous batch

    # -- host -> device, on the copy stream
    def stage(self, v_host, l_host=None):
        """Starts the upload of a batch (pinned uint8 / float32 host tensors).  Call it one batch ahead."""
        main = torch.cuda.current_stream(self.dev)
        with torch.cuda.stream(self.copy_stream):
            v = v_host.to(self.dev, non_blocking=True)
            l = l_host.to(self.dev, non_blocking=True) if l_host is not None else None
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        v.record_stream(main)
        if l is not None:
            l.record_stream(main)
        self._staged = (v, l, ev)

    # -- forward on the current stream, NMS on the nms stream
    def submit(self, v=None, l=None):
        """Runs the forward of a batch (device tensors, or the batch uploaded by `stage`) and queues its NMS.
        Returns the (out [B, max_num, 6], counts [B]) device tensors of the PREVIOUS batch, or None for the first call."""
        main = torch.cuda.current_stream(self.dev)
        if v is None:
            v, l, ev = self._staged
            main.wait_event(ev)
            self._staged = None
        with torch.no_grad():
            io, _ = self.model(v, l) if self.dual else self.model(v)
        done = torch.cuda.Event()
        done.record(main)
        prev = self._pending
        self.nms_stream.wait_event(done)
        io.record_stream(self.nms_stream)
        with torch.cuda.stream(self.nms_stream):
            out, counts = nms_raw(io, *self.args)
            if self.to_host:
                if not self._host_ring or self._host_ring[0][0].shape != out.shape:
                    self._host_ring = [(torch.empty(out.shape, dtype=out.dtype, pin_memory=True),
                                        torch.empty(counts.shape, dtype=counts.dtype, pin_memory=True)) for _ in range(3)]
                h_out, h_cnt = self._host_ring[self._host_slot]
                self._host_slot = (self._host_slot + 1) % len(self._host_ring)
                h_out.copy_(out, non_blocking=True)
                h_cnt.copy_(counts, non_blocking=True)
                out.record_stream(self.nms_stream)
                counts.record_stream(self.nms_stream)
                out, counts = h_out, h_cnt
            fin = torch.cuda.Event()
            fin.record(self.nms_stream)
        self._pending = (out, counts, fin)
        return self._collect(prev)

    def _collect(self, item):
        if item is None:
            return None
        out, counts, fin = item
        if self.to_host:
            fin.synchronize()          # NMS + read-back of the PREVIOUS batch: finished long ago, the forward is not waited for
            return out, counts
        main = torch.cuda.current_stream(self.dev)
        main.wait_event(fin)
        out.record_stream(main)
        counts.record_stream(main)
        return out, counts

    def flush(self):
        item, self._pending = self._pending, None
        return self._collect(item)
