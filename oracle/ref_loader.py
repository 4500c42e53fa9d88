"""Load the reference's OWN class definitions (audiogan.py:1-552) under py3 / torch 2.x.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Works only where
``/root/reference`` exists (this container) -- used by ``oracle/make_golden.py`` to
generate fixtures and by the CPU tests to validate ``oracle/restated.py``.

Recipe:
  1. stub the modules the file imports but the image lacks (tensorflow, librosa,
     matplotlib, PIL, timer, dataset);
  2. exec lines 1-552 (helpers + Generator / Discriminator / Embedder) -- the
     remainder of the file is py2 script body (argparse, HDF5, ``print x``);
  3. four shims forced by py2 / torch<=0.3 semantics:
       tovar          audiogan.py:94-97   drops the unconditional ``.cuda()``
       div_roundup    audiogan.py:172-173 py2 integer ``/`` -> ``//``
       multinomial    audiogan.py:450     no-arg form -> num_samples=1
       int tensor /   audiogan.py:533     LongTensor ``/`` was floor division
"""
import contextlib
import importlib.machinery
import os
import sys
import types

import numpy as NP
import torch as T

REFERENCE_FILE = os.environ.get("AUDIOGAN_REFERENCE", "/root/reference/audiogan.py")
_N_DEF_LINES = 552
_STUBS = ["tensorflow", "librosa", "librosa.feature", "matplotlib", "matplotlib.pyplot",
          "timer", "dataset", "PIL", "PIL.Image"]


def available():
    return os.path.exists(REFERENCE_FILE)


@contextlib.contextmanager
def _stubbed_modules():
    saved = {}
    for name in _STUBS:
        saved[name] = sys.modules.get(name)
        m = types.ModuleType(name)
        m.__spec__ = importlib.machinery.ModuleSpec(name, None)
        sys.modules[name] = m
    sys.modules["matplotlib"].use = lambda *a, **k: None
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["timer"].Timer = object
    sys.modules["librosa"].feature = sys.modules["librosa.feature"]
    sys.modules["PIL"].Image = sys.modules["PIL.Image"]
    try:
        yield
    finally:
        for name, m in saved.items():
            if m is None:
                sys.modules.pop(name, None)
            else:
                sys.modules[name] = m


@contextlib.contextmanager
def py2_tensor_semantics():
    """torch<=0.3 semantics the reference relies on, active only inside the block."""
    orig_div = T.Tensor.__truediv__
    orig_mn = T.Tensor.multinomial

    def _div(self, other):
        int_self = not (self.is_floating_point() or self.is_complex())
        int_other = (isinstance(other, int) and not isinstance(other, bool)) or (
            isinstance(other, T.Tensor) and not (other.is_floating_point() or other.is_complex()))
        if int_self and int_other:
            return T.div(self, other, rounding_mode="floor")
        return orig_div(self, other)

    def _mn(self, num_samples=1, replacement=False, *, generator=None):
        return orig_mn(self, num_samples, replacement, generator=generator)

    T.Tensor.__truediv__ = _div
    T.Tensor.multinomial = _mn
    try:
        yield
    finally:
        T.Tensor.__truediv__ = orig_div
        T.Tensor.multinomial = orig_mn


_NS = None


def load():
    """Returns the namespace holding the reference's Generator, Discriminator, helpers."""
    global _NS
    if _NS is not None:
        return _NS
    if not available():
        raise FileNotFoundError(REFERENCE_FILE)
    with open(REFERENCE_FILE) as f:
        lines = f.readlines()
    src = "".join(lines[:_N_DEF_LINES])
    ns = {"__name__": "audiogan_reference_defs"}
    with _stubbed_modules():
        exec(compile(src, REFERENCE_FILE, "exec"), ns)

    def tovar(*arrs):                                  # audiogan.py:94-97 without .cuda()
        ts = [(T.Tensor(a.astype("float32")) if isinstance(a, NP.ndarray) else a) for a in arrs]
        return ts[0] if len(ts) == 1 else ts

    ns["tovar"] = tovar
    ns["div_roundup"] = lambda x, d: (x + d - 1) // d   # audiogan.py:172-173 (py2 int /)
    ns["roundup"] = lambda x, d: (x + d - 1) // d * d
    _NS = ns
    return ns


def pin_stopper(g, value=30.0):
    """Pin the stop head: bias = g*sign(v) = -value -> never stops."""
    with T.no_grad():
        g.stopper.module.bias_g.fill_(value)
        g.stopper.module.bias_v.fill_(-1.0)
