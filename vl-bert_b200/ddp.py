"""Data-parallel gradient exchange for the encoder (the DDP step of the reference's common/trainer.py:115-127 with
pretrain/function/train.py:90): one process per GPU, the batch is sharded, gradients are averaged with NCCL.

Instead of torch's bucketed reducer (which only sees the encoder's gradients when the single fused autograd node
returns), the per-layer flat gradient buffer is all-reduced asynchronously the moment that layer's backward kernels
are enqueued, so the exchange of layer l overlaps the dgrad/wgrad kernels of layers l-1 .. 0; all reductions are
complete (stream-ordered) before the gradients are handed back to autograd.

Wire format: on NCCL the gradients travel as bf16 (the reference's FP16 path all-reduces half-precision gradients too:
apex DDP, pretrain/function/train.py:353-354): the fp32 buffer is cast into a persistent bf16 staging buffer, that buffer is
all-reduced (AVG), and the result is cast back over the fp32 gradients `pipeline_depth` launches later, by which time the
collective has long finished -- half the bytes on NVLink and half the time NCCL's channel CTAs compete with the
persistent GEMM grids.  `wire_dtype=None` keeps fp32 (gloo in the CPU tests, or VLB_DDP_WIRE=fp32).
On CUDA the two casts and the wait for the collective run on a side stream that forks from the compute stream at launch() and
joins it again in drain(): the 24 cast kernels of a step (~10 us each) no longer sit between the backward kernels of the layers
below (VLB_DDP_SIDE_STREAM=0 puts them back on the compute stream).  The fork / join are ordinary stream dependencies, so a
CUDA-graph capture of the step records them like everything else.
Works with any torch.distributed backend."""
import os

import torch
import torch.distributed as dist

from . import _lib


def _cast(src, dst):
    """dst <- src across fp32 / bf16 (library cast kernels on CUDA, torch on the CPU)"""
    if src.is_cuda and src.is_contiguous() and dst.is_contiguous() and src.data_ptr() % 16 == 0 and dst.data_ptr() % 16 == 0:
        st = torch.cuda.current_stream().cuda_stream
        lib = _lib.lib()
        if src.dtype == torch.float32 and dst.dtype == torch.bfloat16:
            _lib.check(lib.vlb_cast_f32_to_bf16(src.data_ptr(), dst.data_ptr(), src.numel(), st))
            return
        if src.dtype == torch.bfloat16 and dst.dtype == torch.float32:
            _lib.check(lib.vlb_cast_bf16_to_f32(src.data_ptr(), dst.data_ptr(), src.numel(), st))
            return
    dst.copy_(src)


class LayerGradReducer(object):
    def __init__(self, group=None, wire_dtype="auto", pipeline_depth=2):
        self.group = group
        self.world = dist.get_world_size(group)
        self.avg = dist.get_backend(group) == "nccl"
        if wire_dtype == "auto":
            env = os.environ.get("VLB_DDP_WIRE", "bf16" if self.avg else "fp32")
            wire_dtype = torch.bfloat16 if env == "bf16" else None
        self.wire_dtype = wire_dtype
        self.depth = max(0, int(pipeline_depth))
        self.pending = []
        self._staging = {}
        self.side_stream = os.environ.get("VLB_DDP_SIDE_STREAM", "1") != "0"
        self._side = {}
        self._forked = set()

    def _stage(self, flat):
        """persistent wire-format buffer for this gradient buffer (keyed by address: static under CUDA graphs)"""
        key = (flat.data_ptr(), flat.numel())
        buf = self._staging.get(key)
        if buf is None:
            if len(self._staging) > 256:
                self._staging.clear()
            buf = torch.empty(flat.numel(), dtype=self.wire_dtype, device=flat.device)
            self._staging[key] = buf
        return buf

    def launch(self, flat):
        """flat: contiguous fp32 gradient buffer of one layer; reduced in place (visible after drain())."""
        op = dist.ReduceOp.AVG if self.avg else dist.ReduceOp.SUM
        if self.side_stream and flat.is_cuda and self.avg and flat.is_contiguous():
            # cast -> all-reduce -> cast back, all off the compute stream; drain() joins
            cur = torch.cuda.current_stream(flat.device)
            side = self._side.get(flat.device)
            if side is None:
                side = self._side[flat.device] = torch.cuda.Stream(device=flat.device)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                if self.wire_dtype is not None and flat.dtype == torch.float32:
                    wire = self._stage(flat)
                    _cast(flat.view(-1), wire)
                    dist.all_reduce(wire, op=op, group=self.group, async_op=True).wait()
                    _cast(wire, flat.view(-1))
                else:
                    dist.all_reduce(flat, op=op, group=self.group, async_op=True).wait()
            self._forked.add(flat.device)
            return
        if self.wire_dtype is not None and flat.dtype == torch.float32 and flat.is_contiguous():
            wire = self._stage(flat)
            _cast(flat.view(-1), wire)
            h = dist.all_reduce(wire, op=op, group=self.group, async_op=True)
            self.pending.append((h, flat, wire))
        else:
            h = dist.all_reduce(flat, op=op, group=self.group, async_op=True)
            self.pending.append((h, flat, None))
        while len(self.pending) > self.depth:       # finish the collective launched `depth` launches ago
            self._finish(self.pending.pop(0))

    def _finish(self, item):
        h, flat, wire = item
        h.wait()
        if wire is not None:
            _cast(wire, flat.view(-1))
        if not self.avg:
            flat.mul_(1.0 / self.world)

    def drain(self):
        while self.pending:
            self._finish(self.pending.pop(0))
        for dev in self._forked:
            torch.cuda.current_stream(dev).wait_stream(self._side[dev])
        self._forked.clear()

    def reduce_params(self, params, coalesce_below=1 << 20):
        """Average the .grad of parameters that are not covered by launch() (embeddings, pooler, heads).
        Large gradients (the dense [vocab, H] word-embedding gradient) are reduced in place; the many small ones are
        coalesced into one flat buffer (one collective instead of dozens)."""
        gs = [p.grad for p in params if p.grad is not None]
        if not gs:
            return
        small = [g for g in gs if g.numel() < coalesce_below or not g.is_contiguous()]
        flat = None
        if small:                       # the small ones first: they are ready and tiny
            flat = torch.cat([g.reshape(-1) for g in small])
            self.launch(flat)
        for g in gs:
            if g.numel() >= coalesce_below and g.is_contiguous():
                self.launch(g)
        self.drain()
        if flat is not None:
            o = 0
            for g in small:
                g.copy_(flat[o:o + g.numel()].view_as(g))
                o += g.numel()


def attach(model, group=None, **kw):
    """Enable the overlapped gradient exchange on a vlbert_b200.VisualLinguisticBert; returns the reducer."""
    r = LayerGradReducer(group, **kw)
    model._grad_reducer = r
    return r
