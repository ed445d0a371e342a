"""Torch7 ``nn.Sequential`` semantics restated on CPU (oracle; test infrastructure only).

A net is built from a plain layer-spec list (dicts), the same description the product's
host API turns into ``dcgansr_layer`` records, e.g.::

    [dict(kind="fullconv", cin=3, cout=96, k=4, s=2, p=1), dict(kind="bn", c=96),
     dict(kind="relu"), ...]

What is restated (reference call sites in parentheses):

* every module caches its ``output`` at ``forward`` and the *last forward wins*
  (``netD.output`` reuse, train.lua:265);
* ``backward`` = ``updateGradInput`` + ``accGradParameters`` (grads accumulate, ``+=``);
  ``updateGradInput`` alone walks the graph with the *current* weights and the *cached*
  activations / BN ``save_mean``/``save_std`` (train.lua:268, fact F5);
* ``getParameters()`` order: per module ``weight`` then ``bias`` (train.lua:202-203);
* BN is always in training mode (``:evaluate()`` is never called, fact F7); running
  statistics are maintained.

PARITY UNPINNED: see ``oracle/__init__.py``.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops

_ACT = {"relu": ops.ACT_RELU, "lrelu": ops.ACT_LRELU, "tanh": ops.ACT_TANH, "sigmoid": ops.ACT_SIGMOID}


class Module:
    kind = "?"

    def __init__(self):
        self.output = None
        self.input = None

    def params(self):  # [(name, tensor, grad)]
        return []


class Conv(Module):
    def __init__(self, spec, dtype, full):
        super().__init__()
        self.full = full
        self.cin, self.cout = spec["cin"], spec["cout"]
        self.k, self.s, self.p = spec["k"], spec.get("s", 1), spec.get("p", 0)
        self.adj = spec.get("adj", 0)
        shape = (self.cin, self.cout, self.k, self.k) if full else (self.cout, self.cin, self.k, self.k)
        self.weight = torch.zeros(shape, dtype=dtype)
        self.gradWeight = torch.zeros(shape, dtype=dtype)

    def forward(self, x):
        self.input = x
        if self.full:
            self.output = ops.fullconv2d_fwd(x, self.weight, self.s, self.p, self.adj)
        else:
            self.output = ops.conv2d_fwd(x, self.weight, self.s, self.p)
        return self.output

    def update_grad_input(self, x, dy):
        if self.full:
            return ops.fullconv2d_dgrad(dy, self.weight, self.s, self.p)
        return ops.conv2d_dgrad(dy, self.weight, x.shape, self.s, self.p)

    def acc_grad_parameters(self, x, dy):
        if self.full:
            self.gradWeight += ops.fullconv2d_wgrad(x, dy, self.weight.shape, self.s, self.p)
        else:
            self.gradWeight += ops.conv2d_wgrad(x, dy, self.weight.shape, self.s, self.p)

    def params(self):
        return [("weight", self.weight, self.gradWeight)]


class BN(Module):
    def __init__(self, spec, dtype):
        super().__init__()
        c = spec["c"]
        self.eps, self.momentum = spec.get("eps", 1e-5), spec.get("momentum", 0.1)
        self.weight = torch.ones(c, dtype=dtype)
        self.bias = torch.zeros(c, dtype=dtype)
        self.gradWeight = torch.zeros(c, dtype=dtype)
        self.gradBias = torch.zeros(c, dtype=dtype)
        self.running_mean = torch.zeros(c, dtype=dtype)
        self.running_var = torch.ones(c, dtype=dtype)
        self.save_mean = None
        self.save_std = None  # holds invstd, as in Torch7

    def forward(self, x):
        self.input = x
        y, m, istd, rm, rv = ops.bn_fwd_train(x, self.weight, self.bias, self.running_mean,
                                              self.running_var, self.eps, self.momentum)
        self.save_mean, self.save_std = m, istd
        self.running_mean, self.running_var = rm, rv
        self.output = y
        return y

    def update_grad_input(self, x, dy):
        dx, self._dg, self._db = ops.bn_bwd(x, dy, self.weight, self.save_mean, self.save_std)
        return dx

    def acc_grad_parameters(self, x, dy):
        _, dg, db = ops.bn_bwd(x, dy, self.weight, self.save_mean, self.save_std)
        self.gradWeight += dg
        self.gradBias += db

    def params(self):
        return [("weight", self.weight, self.gradWeight), ("bias", self.bias, self.gradBias)]


class Act(Module):
    def __init__(self, kind, negval=0.2):
        super().__init__()
        self.akind, self.negval = _ACT[kind], negval

    def forward(self, x):
        self.input = x
        self.output = ops.act_fwd(x, self.akind, self.negval)
        return self.output

    def update_grad_input(self, x, dy):
        return ops.act_bwd(self.output, dy, self.akind, self.negval)

    def acc_grad_parameters(self, x, dy):
        pass


class UpNearest(Module):
    def __init__(self, scale=2):
        super().__init__()
        self.scale = scale

    def forward(self, x):
        self.input = x
        self.output = ops.upnearest_fwd(x, self.scale)
        return self.output

    def update_grad_input(self, x, dy):
        return ops.upnearest_bwd(dy, self.scale)

    def acc_grad_parameters(self, x, dy):
        pass


class View(Module):
    """nn.View(1):setNumInputDims(3): B x 1 x h x w -> (B*h*w) x 1 (train.lua:136)."""

    def forward(self, x):
        self.input = x
        self.output = x.reshape(-1, 1)
        return self.output

    def update_grad_input(self, x, dy):
        return dy.reshape(x.shape)

    def acc_grad_parameters(self, x, dy):
        pass


def _make(spec, dtype):
    k = spec["kind"]
    if k == "conv":
        return Conv(spec, dtype, full=False)
    if k == "fullconv":
        return Conv(spec, dtype, full=True)
    if k == "bn":
        return BN(spec, dtype)
    if k in _ACT:
        return Act(k, spec.get("negval", 0.2))
    if k == "upnearest":
        return UpNearest(spec.get("scale", 2))
    if k == "view":
        return View()
    raise ValueError(k)


class Sequential:
    def __init__(self, specs, dtype=torch.float64):
        self.dtype = dtype
        self.specs = [dict(s) for s in specs]
        self.modules = [_make(s, dtype) for s in specs]
        self.output = None

    # -- parameters --------------------------------------------------------------------
    def param_list(self):
        out = []
        for m in self.modules:
            out.extend(m.params())
        return out

    def num_params(self):
        return sum(p.numel() for _, p, _ in self.param_list())

    def get_flat_params(self):
        return torch.cat([p.reshape(-1) for _, p, _ in self.param_list()])

    def get_flat_grads(self):
        return torch.cat([g.reshape(-1) for _, _, g in self.param_list()])

    def set_flat_params(self, flat):
        flat = torch.as_tensor(flat, dtype=self.dtype)
        off = 0
        for _, p, _ in self.param_list():
            n = p.numel()
            p.copy_(flat[off:off + n].reshape(p.shape))
            off += n
        assert off == flat.numel()

    def zero_grad_parameters(self):
        for _, _, g in self.param_list():
            g.zero_()

    def bn_modules(self):
        return [m for m in self.modules if isinstance(m, BN)]

    # -- Torch7 walk ---------------------------------------------------------------------
    def forward(self, x):
        cur = x
        for m in self.modules:
            cur = m.forward(cur)
        self.output = cur
        return cur

    def _walk_back(self, x, dy, acc):
        cur = dy
        for i in range(len(self.modules) - 1, -1, -1):
            m = self.modules[i]
            inp = self.modules[i - 1].output if i > 0 else x
            if acc:
                m.acc_grad_parameters(inp, cur)
            cur = m.update_grad_input(inp, cur)
        return cur

    def backward(self, x, dy):
        return self._walk_back(x, dy, acc=True)

    def update_grad_input(self, x, dy):
        return self._walk_back(x, dy, acc=False)

    def kink_margin(self):
        """Conditioning probe for the parity tests: the smallest |pre-activation| / max|pre-activation| over all
        ReLU / LeakyReLU inputs of the last forward.  An element within float32 rounding (~1e-7) of the kink may
        take the other branch in ANY float32 implementation (Torch7 included), which changes the gradients at the
        1e-3 level -- a property of the network, not of an implementation."""
        m = float("inf")
        for mod in self.modules:
            if isinstance(mod, Act) and mod.akind in (ops.ACT_RELU, ops.ACT_LRELU) and mod.input is not None:
                a = mod.input.abs()
                mx = float(a.max())
                if mx > 0:
                    m = min(m, float(a.min()) / mx)
        return m


def weights_init(net: Sequential, seed: int):
    """weights_init (train.lua:42-51): conv ~N(0,.02) bias-free, BN gamma~N(1,.02), beta=0.

    The reference's seed is itself random (train.lua:30-32) so parity injects identical
    *weights*; we draw them with numpy Philox, rounded to
    float32 first, so that the float64 oracle, the float32 oracle and the library share bits.
    """
    rng = np.random.Generator(np.random.Philox(seed))
    for m in net.modules:
        if isinstance(m, Conv):
            m.weight.copy_(torch.from_numpy(rng.normal(0.0, 0.02, size=tuple(m.weight.shape)).astype(np.float32)).to(net.dtype))
        elif isinstance(m, BN):
            m.weight.copy_(torch.from_numpy(rng.normal(1.0, 0.02, size=tuple(m.weight.shape)).astype(np.float32)).to(net.dtype))
            m.bias.zero_()
    return net
