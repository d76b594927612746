"""Run the reference's own classes in THIS container (no import, no copy).

The reference scripts cannot be imported (hyphenated names, module-level MNIST
download, missing tensorboardX/seaborn, module-level training loop), so the
class definitions are AST-sliced out of the files where they lie under
/root/reference and exec'd with the module globals they read.
Nothing from the reference is written into this repo: only the OUTPUTS of
running it are stored, as golden vectors (see make_golden.py).

Also provides the noise replay used to drive the reference with known noise:
torch's samplers are patched for the duration of a call so that every draw pops
the next pre-generated tensor from a queue (draw order per layer: SURVEY.md
§3.2).  `torch.bernoulli(p)` pops a uniform `u` and returns `u < p`, which is
the definition of the inclusion / flow masks used everywhere in this repo.

This file is generation-time tooling: it needs /root/reference and is never
imported by tests, bench.py or the product.
"""
import ast
import contextlib
import math
import os
import sys

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

REF_ROOT = os.environ.get("LBBNN_REFERENCE", "/root/reference")

_CLASSES = {"Gaussian", "Bernoulli", "GaussGamma", "BetaBinomial", "BayesianLinear", "BayesianNetwork", "BayesianLayer", "BNN"}


def load_reference_classes(script, flows_module=None, functions=(), **globals_override):
    """exec the ClassDefs (and the named FunctionDefs) of `script` (e.g. 'LBBNN-GP-MF-LRT.py') and return the namespace."""
    path = os.path.join(REF_ROOT, script)
    with open(path) as fh:
        tree = ast.parse(fh.read(), filename=path)
    body = [n for n in tree.body if (isinstance(n, ast.ClassDef) and n.name in _CLASSES) or
            (isinstance(n, ast.FunctionDef) and n.name in functions)]
    module = ast.Module(body=body, type_ignores=[])
    ns = {
        "torch": torch, "nn": nn, "F": F, "math": math, "np": np,
        "DEVICE": torch.device("cpu"),
        "TEMPER_PRIOR": 0.001, "TEMPER": 0.001,
        "SAMPLES": 1, "BATCH_SIZE": 100, "CLASSES": 10, "NUM_BATCHES": 600,
        "Z_FLOW_TYPE": "RNVP", "R_FLOW_TYPE": "RNVP",
    }
    if flows_module is not None:
        if REF_ROOT not in sys.path:
            sys.path.insert(0, REF_ROOT)
        mod = __import__(flows_module)
        ns["PropagateFlow"] = mod.PropagateFlow
        ns["_flows"] = mod
    ns.update(globals_override)
    exec(compile(module, path, "exec"), ns)
    return ns


class NoiseQueue:
    """Pre-generated noise handed to the reference in draw order."""

    def __init__(self, items=()):
        self.items = list(items)
        self.log = []

    def pop(self, kind, shape):
        if not self.items:
            raise RuntimeError(f"noise queue exhausted at draw #{len(self.log)} ({kind} {tuple(shape)})")
        k, t = self.items.pop(0)
        if k != kind or tuple(t.shape) != tuple(shape):
            raise RuntimeError(f"draw #{len(self.log)}: reference asked for {kind}{tuple(shape)}, "
                               f"queue holds {k}{tuple(t.shape)}")
        self.log.append((kind, tuple(shape)))
        return t


class NoiseTape:
    """Records every draw (kind, tensor) the reference makes with the real RNG."""

    def __init__(self):
        self.items = []


@contextlib.contextmanager
def replay(queue):
    """Patch torch samplers so the reference consumes `queue` instead of the RNG."""
    saved = {n: getattr(torch, n) for n in ("randn", "randn_like", "bernoulli", "rand", "normal", "_standard_gamma")}

    def randn(*size, **kw):
        size = kw.pop("size", size)
        if len(size) == 1 and not isinstance(size[0], int):
            size = tuple(size[0])
        return queue.pop("normal", size).clone()

    def randn_like(t, **kw):
        return queue.pop("normal", t.shape).clone()

    def bernoulli(p, *a, **kw):
        u = queue.pop("uniform", p.shape)
        return (u < p).to(p.dtype)

    def rand(*size, **kw):
        size = kw.pop("size", size)
        if len(size) == 1 and not isinstance(size[0], int):
            size = tuple(size[0])
        return queue.pop("uniform", size).clone()

    def normal(mean, std, *a, **kw):
        shape = torch.broadcast_shapes(getattr(mean, "shape", ()), getattr(std, "shape", ()))
        return mean + std * queue.pop("normal", shape)

    class _StdGamma(torch.autograd.Function):
        # keeps the implicit reparameterisation gradient of torch._standard_gamma
        @staticmethod
        def forward(ctx, conc):
            g = queue.pop("gamma", conc.shape).clone()
            ctx.save_for_backward(conc.detach(), g)
            return g

        @staticmethod
        def backward(ctx, grad):
            conc, g = ctx.saved_tensors
            return grad * torch._standard_gamma_grad(conc, g)

    def standard_gamma(conc, *a, **kw):
        return _StdGamma.apply(conc)

    torch.randn, torch.randn_like, torch.bernoulli = randn, randn_like, bernoulli
    torch.rand, torch.normal, torch._standard_gamma = rand, normal, standard_gamma
    import torch.distributions.gamma as _g
    import torch.distributions.relaxed_bernoulli as _rb
    g_saved, rb_saved = _g._standard_gamma, None
    _g._standard_gamma = standard_gamma
    try:
        yield queue
    finally:
        for n, f in saved.items():
            setattr(torch, n, f)
        _g._standard_gamma = g_saved


@contextlib.contextmanager
def record(tape):
    """Let the reference use the real RNG but keep a copy of every draw."""
    saved = {n: getattr(torch, n) for n in ("randn", "randn_like", "bernoulli", "rand", "normal")}

    def wrap(name, kind):
        fn = saved[name]

        def inner(*a, **kw):
            out = fn(*a, **kw)
            tape.items.append((name, tuple(out.shape)))
            return out
        return inner

    for n in saved:
        setattr(torch, n, wrap(n, n))
    try:
        yield tape
    finally:
        for n, f in saved.items():
            setattr(torch, n, f)
