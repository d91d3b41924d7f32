"""CUDA-graph drivers of the two loops around the hot path.

The reference runs `predict_with_uncertainty` as a host loop of S eager forwards (models.py:361-364, ~15 k ATen calls per forward
at cfg 1) and a trainer step as one eager forward + backward (training/trainer.py:242-313).  Here ONE sampled forward (+ the
on-device moment accumulation), or one forward + loss + backward, is captured into a CUDA graph and replayed: the graph's kernels
read the Philox (seed, sample_idx) from a 3-word device buffer (`pno_set_dynamic_noise`), so every replay draws the noise of its
own sample / step without re-recording launch arguments.  Results are bit-identical to the eager loop (same kernels, same order).
"""
import weakref

import torch

from . import _lib
from . import functional as F_


_replayed = [0]          # kernels launched through graph replays (each replay re-runs the launches recorded at capture time)


def replayed_launches():
    """Kernels of this library launched by graph replays so far (pno_launch_count only sees host-side launches)."""
    return _replayed[0]


class NoiseKey:
    """The device-resident (seed lo, seed hi, sample_idx, unused) words the graph's kernels read."""

    def __init__(self, device):
        self.dev = torch.zeros(4, dtype=torch.int32, device=device)

    def set(self, seed, sample_idx):
        with _lib.on_device(self.dev):
            _lib.check(_lib.load().pno_noise_key_write(self.dev.data_ptr(), int(seed) & (2 ** 64 - 1), int(sample_idx) & 0xFFFFFFFF,
                                                       _lib.stream_ptr(self.dev.device)), "pno_noise_key_write")

    def advance(self, delta=1):
        with _lib.on_device(self.dev):
            _lib.check(_lib.load().pno_noise_key_advance(self.dev.data_ptr(), int(delta), _lib.stream_ptr(self.dev.device)),
                       "pno_noise_key_advance")

    def __enter__(self):
        _lib.check(_lib.load().pno_set_dynamic_noise(self.dev.data_ptr()), "pno_set_dynamic_noise")
        return self

    def __exit__(self, *exc):
        _lib.check(_lib.load().pno_set_dynamic_noise(None), "pno_set_dynamic_noise")
        return False


def _param_signature(model):
    return tuple((p.data_ptr(), tuple(p.shape)) for p in model.parameters())


class MCForwardGraph:
    """One captured `y = model(x, sample=True); s1 += y; s2 += y^2; sample_idx += 1` for a fixed input shape."""

    def __init__(self, model, x):
        self.sig = (tuple(x.shape), x.device, model.engine, _param_signature(model))
        dev = x.device
        self.x = torch.empty_like(x)
        self.x.copy_(x)
        self.key = NoiseKey(dev)
        self.moments = torch.zeros((2, x.shape[0], 1) + tuple(x.shape[2:]), dtype=torch.float64, device=dev)
        self.stream = torch.cuda.Stream(device=dev)
        self.graph = torch.cuda.CUDAGraph()
        self.stream.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(self.stream), self.key:
            self.key.set(0, 0)
            # eager warm-up on the capture stream: plans, per-stream scratch and function attributes are created here, and any
            # argument error is raised here, outside the capture
            self._body(model)
            self.stream.synchronize()
            n0 = _lib.load().pno_launch_count()
            with torch.cuda.graph(self.graph, stream=self.stream):
                self._body(model)
            self.launches = _lib.load().pno_launch_count() - n0          # library kernels recorded in the graph
        torch.cuda.current_stream(dev).wait_stream(self.stream)

    def _body(self, model):
        with F_.noise_scope(0, 0, 0):                 # by-value (seed, sample) are ignored under the dynamic key
            y = model.forward(self.x, sample=True)
        F_.mc_accumulate(y, self.moments[0], self.moments[1])
        self.key.advance(1)

    def matches(self, model, x):
        return self.sig == (tuple(x.shape), x.device, model.engine, _param_signature(model))

    def run(self, x, seed, first_sample, count):
        """Accumulate samples [first_sample, first_sample + count) of input x; returns the (sum, sum of squares) buffers."""
        self.x.copy_(x)
        self.moments.zero_()
        self.key.set(seed, first_sample)
        for _ in range(count):
            self.graph.replay()
        _replayed[0] += count * self.launches
        return self.moments[0], self.moments[1]


_MC_GRAPHS = weakref.WeakKeyDictionary()      # model -> [MCForwardGraph, ...] (kept off the module: graphs do not pickle / deepcopy)


def mc_graph_for(model, x):
    """The captured sampled forward of `model` for inputs shaped like x (at most two per model: a graph holds its activations)."""
    graphs = _MC_GRAPHS.setdefault(model, [])
    for g in graphs:
        if g.matches(model, x):
            return g
    g = MCForwardGraph(model, x.contiguous())
    graphs.insert(0, g)
    del graphs[2:]
    return g


def drop_graphs(model):
    _MC_GRAPHS.pop(model, None)


class TrainStepGraph:
    """One captured `zero grads; outputs = model(x, sample=True, return_kl=True); loss = loss_fn(...); loss.backward()` for fixed
    input / target shapes.  The optimiser step stays outside (its hyper-parameters and step count are host values)."""

    def __init__(self, model, loss_fn, x, y, step_fn):
        dev = x.device
        self.sig = (tuple(x.shape), tuple(y.shape), dev, model.engine, _param_signature(model))
        self.x, self.y = torch.empty_like(x), torch.empty_like(y)
        self.x.copy_(x)
        self.y.copy_(y)
        self.key = NoiseKey(dev)
        self.stream = torch.cuda.Stream(device=dev)
        self.graph = torch.cuda.CUDAGraph()
        self.step_fn = step_fn
        self.out = None
        params = [p for p in model.parameters() if p.requires_grad]
        saved = [None if p.grad is None else p.grad for p in params]      # the caller's gradients survive construction
        self.stream.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(self.stream), self.key:
            self.key.set(0, 0)
            for _ in range(2):                                           # eager warm-up (autograd allocates the .grad buffers)
                for p in params:
                    p.grad = None
                step_fn(self.x, self.y)
            self.stream.synchronize()
            for p in params:
                p.grad = None
            n0 = _lib.load().pno_launch_count()
            with torch.cuda.graph(self.graph, stream=self.stream):
                self.out = step_fn(self.x, self.y)
            self.launches = _lib.load().pno_launch_count() - n0
        torch.cuda.current_stream(dev).wait_stream(self.stream)
        self.grads = [p.grad for p in params]                             # static gradient buffers owned by the graph's pool
        self.params = params
        for p, g in zip(params, saved):
            p.grad = g

    def matches(self, model, x, y):
        return self.sig == (tuple(x.shape), tuple(y.shape), x.device, model.engine, _param_signature(model))

    def run(self, x, y, seed, step):
        """Replays the step on (x, y) with the noise of (seed, step); .grad of every parameter is the graph's static buffer."""
        self.x.copy_(x, non_blocking=True)
        self.y.copy_(y, non_blocking=True)
        self.key.set(seed, step)
        self.graph.replay()
        _replayed[0] += self.launches
        for p, g in zip(self.params, self.grads):
            p.grad = g
        return self.out
