"""CUDA-graph capture of a whole forward+backward step of the hot path.

The step is ~270 kernel launches of 10-70 us each; launched eagerly they leave ~1 ms of gaps per step (Python + driver
launch latency).  Shapes are static in training (fixed batch, `max_length_hint`), so the step can be captured once and
replayed: `GraphedStep(model, loss_fn, example_inputs)` warms up, captures `loss = loss_fn(model, *inputs);
loss.backward()` into a torch.cuda.CUDAGraph with static input buffers, and `__call__(*inputs)` copies the new batch into
those buffers and replays.  Parameter gradients live in static buffers (`p.grad`) that the replay overwrites.
"""
import torch


class GraphedStep(object):
    def __init__(self, model, loss_fn, example_inputs, warmup=3, reducer_params=None):
        self.model = model
        self.loss_fn = loss_fn
        self.static_inputs = [t.clone() if torch.is_tensor(t) else t for t in example_inputs]
        self.reducer = getattr(model, "_grad_reducer", None)
        self.reducer_params = reducer_params
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self._eager()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        model.zero_grad(set_to_none=True)
        from . import functional as _F
        _F._CAPTURE_HAS_BACKWARD = True        # side-stream work forked in the forward is joined by the backward of this capture
        try:
            with torch.cuda.graph(self.graph):
                self.static_loss = self._eager(zero=False)
        finally:
            _F._CAPTURE_HAS_BACKWARD = False
        torch.cuda.synchronize()
        # gradients produced by a replay live in these buffers
        self.static_grads = [(p, p.grad) for p in model.parameters() if p.grad is not None]

    def _eager(self, zero=True):
        if zero:
            self.model.zero_grad(set_to_none=True)
        loss = self.loss_fn(self.model, *self.static_inputs)
        loss.backward()
        if self.reducer is not None and self.reducer_params is not None:
            self.reducer.reduce_params(self.reducer_params)
        return loss.detach()

    def load_inputs(self, inputs, non_blocking=True):
        for dst, src in zip(self.static_inputs, inputs):
            if torch.is_tensor(dst):
                dst.copy_(src, non_blocking=non_blocking)

    def __call__(self, *inputs):
        if inputs:
            self.load_inputs(inputs)
        self.graph.replay()
        for p, g in self.static_grads:
            if p.grad is not g:
                p.grad = g
        return self.static_loss
