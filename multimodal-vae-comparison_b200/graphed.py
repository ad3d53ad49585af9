"""CUDA-graph capture of a whole plugin step: ``model.objective(batch)["loss"].backward()``.

The drop-in classes issue ~60 launches per step (stand-in encoders / decoders in torch + the C-ABI kernels); eager, the
step is bound by launch latency and Python.  For a fixed batch signature (same modalities, shapes, dtypes, mask
presence) the whole chain -- encoders, latent kernels, decoders, likelihood kernels, objective, backward of all of it --
is captured once and replayed as ONE graph launch; inputs are copied into static buffers, gradients land in static
``.grad`` tensors, noise is drawn inside the graph (torch's CUDA generator is capture aware: every replay advances the
Philox offset, so replays see fresh noise exactly like eager steps).

    g = GraphedObjective(model, example_batch)
    for batch in loader:
        out = g.step(batch)          # {"loss", "kld", "reconstruction_loss"} -- static tensors, valid until the next step
        optimizer.step()             # gradients are in p.grad (re-attached by step(); do not zero them: replays overwrite)

Everything the objective does on the host while capturing (subset tables, MoPoE chunk maps, plan decisions) is frozen
into the graph; a batch with a different signature needs a new GraphedObjective.  The objective must be free of host
synchronisation -- true for every model / objective pair of this package.
"""
from typing import Dict

import torch


def _as_static(v):
    """Batch entries the graph must see through static buffers: tensors, and the reference's list-of-tensors form of
    ``data`` (objectives._prep stacks it) -- stacked once here so that replays see fresh values."""
    if isinstance(v, (list, tuple)) and len(v) and all(torch.is_tensor(x) for x in v):
        return torch.stack(list(v))
    return v


def _signature(batch: Dict[str, dict]):
    sig = []
    for mod in sorted(batch):
        for k in sorted(batch[mod]):
            v = _as_static(batch[mod][k])
            sig.append((mod, k, (tuple(v.shape), v.dtype) if torch.is_tensor(v) else v))
    return tuple(sig)


class GraphedObjective:
    def __init__(self, model, batch: Dict[str, dict], warmup: int = 3, backward: bool = True):
        dev = next(model.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("GraphedObjective needs a model on a CUDA device (there is no CPU path)")
        self.model, self.backward = model, backward
        self.signature = _signature(batch)
        self.static = {mod: {k: (_as_static(v).detach().to(dev).clone() if torch.is_tensor(_as_static(v)) else v)
                             for k, v in entry.items()} for mod, entry in batch.items()}
        self.params = [p for p in model.parameters() if p.requires_grad]
        self._one = None
        if backward:
            from . import ops
            self._one = ops.mark_unit_grad(torch.ones((), device=dev))
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):  # lazy initialisation (cuBLAS handles, workspaces, autotuning) outside the capture
                self._eager()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        for p in self.params:
            p.grad = None
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = self._eager()
        self.grads = [p.grad for p in self.params]

    def _eager(self):
        for p in self.params:
            p.grad = None
        out = self.model.objective(self.static)
        if self.backward:
            out["loss"].backward(self._one)
        return out

    def step(self, batch: Dict[str, dict]):
        if _signature(batch) != self.signature:
            raise RuntimeError("GraphedObjective: the batch signature changed (shapes / dtypes / mask presence); "
                               "capture a new graph for it")
        for mod, entry in batch.items():
            for k, v in entry.items():
                v = _as_static(v)
                if torch.is_tensor(v):
                    self.static[mod][k].copy_(v, non_blocking=True)
        self.graph.replay()
        for p, g in zip(self.params, self.grads):
            p.grad = g
        return self.out

    def close(self):
        """Destroy the graph (needed before dist.destroy_process_group() if a collective was captured)."""
        if self.graph is not None:
            self.graph.reset()
            self.graph = None
        if self._one is not None:
            from . import ops
            ops.unmark_unit_grad(self._one)  # the registry would keep it (and its address) alive forever otherwise
            self._one = None
