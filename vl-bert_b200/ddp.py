"""Data-parallel gradient exchange for the encoder (the DDP step of the reference's common/trainer.py:115-127 with
pretrain/function/train.py:90): one process per GPU, the batch is sharded, gradients are averaged with NCCL.

Instead of torch's bucketed reducer (which only sees the encoder's gradients when the single fused autograd node
returns), the per-layer flat gradient buffer is all-reduced asynchronously the moment that layer's backward kernels
are enqueued, so the exchange of layer l overlaps the dgrad/wgrad kernels of layers l-1 .. 0; all reductions are
complete (stream-ordered) before the gradients are handed back to autograd.  Works with any torch.distributed backend
(NCCL on GPUs; gloo in the CPU tests)."""
import torch
import torch.distributed as dist


class LayerGradReducer(object):
    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.avg = dist.get_backend(group) == "nccl"
        self.pending = []

    def launch(self, flat):
        """flat: contiguous gradient buffer of one layer; reduced in place."""
        op = dist.ReduceOp.AVG if self.avg else dist.ReduceOp.SUM
        self.pending.append((dist.all_reduce(flat, op=op, group=self.group, async_op=True), flat))

    def drain(self):
        for h, flat in self.pending:
            h.wait()
            if not self.avg:
                flat.mul_(1.0 / self.world)
        self.pending = []

    def reduce_params(self, params, coalesce_below=1 << 20):
        """Average the .grad of parameters that are not covered by launch() (embeddings, pooler, heads).
        Large gradients (the dense [vocab, H] word-embedding gradient) are reduced in place; the many small ones are
        coalesced into one flat buffer (one collective instead of dozens)."""
        gs = [p.grad for p in params if p.grad is not None]
        if not gs:
            return
        small = [g for g in gs if g.numel() < coalesce_below or not g.is_contiguous()]
        for g in gs:
            if g.numel() >= coalesce_below and g.is_contiguous():
                self.launch(g)
        flat = None
        if small:
            flat = torch.cat([g.reshape(-1) for g in small])
            self.launch(flat)
        self.drain()
        if flat is not None:
            o = 0
            for g in small:
                g.copy_(flat[o:o + g.numel()].view_as(g))
                o += g.numel()


def attach(model, group=None):
    """Enable the overlapped gradient exchange on a vlbert_b200.VisualLinguisticBert; returns the reducer."""
    r = LayerGradReducer(group)
    model._grad_reducer = r
    return r
