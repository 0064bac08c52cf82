"""Load the UNMODIFIED reference hot-path modules from /root/reference.

TEST INFRASTRUCTURE ONLY, and build-container only: /root/reference does not
exist on the GPU box, so nothing in ``-m gpu`` tests, ``smoke()`` or
``bench.py`` may call this.  It is used by ``oracle/make_golden.py`` (to write
``tests/golden``) and by the container-only test that compares the oracle
with the live reference.

The reference's package ``__init__`` files import audio/plotting packages
that are not installed (zounds, conjure, matplotlib, ...), so stub *packages*
named ``modules`` / ``util`` / ``zounds`` are registered first; the reference
source files themselves are imported untouched.
"""
from __future__ import annotations

import ast
import importlib
import os
import sys
import types

REFERENCE_ROOT = "/root/reference"


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "modules"))


def load():
    """Returns a namespace with ``matchingpursuit``, ``conv``, ``fft``,
    ``sparse``, ``decompose``, ``normalization``, ``multibanddict`` and
    ``MatchingPursuit`` (the class body of mp.py:32-67 executed from the
    reference file itself)."""
    if not available():
        raise RuntimeError("/root/reference is not present on this machine")
    import torch

    sys.dont_write_bytecode = True
    saved = {name: sys.modules.get(name) for name in ("modules", "util", "zounds")}
    pkg = types.ModuleType("modules"); pkg.__path__ = [os.path.join(REFERENCE_ROOT, "modules")]
    util = types.ModuleType("util"); util.__path__ = [os.path.join(REFERENCE_ROOT, "util")]
    util.device = torch.device("cpu")
    zounds = types.ModuleType("zounds"); zounds.SampleRate = object; zounds.SR22050 = lambda: 22050
    sys.modules["modules"], sys.modules["util"], sys.modules["zounds"] = pkg, util, zounds
    ns = types.SimpleNamespace()
    for name in ("normalization", "conv", "fft", "sparse", "decompose", "matchingpursuit", "multibanddict"):
        setattr(ns, name, importlib.import_module("modules." + name))

    # mp.py cannot be imported (matplotlib Qt backend, conjure, data); run the
    # class definition it contains, verbatim, against the reference's own
    # fft_convolve (modules/fft.py is arithmetically identical to the
    # modules/transfer.py copy mp.py imports) and sparsify2.
    with open(os.path.join(REFERENCE_ROOT, "mp.py")) as fh:
        tree = ast.parse(fh.read())
    cls = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "MatchingPursuit"]
    env = {"torch": torch, "nn": torch.nn, "fft_convolve": ns.fft.fft_convolve,
           "sparsify2": ns.sparse.sparsify2}
    exec(compile(ast.Module(body=cls, type_ignores=[]), "mp.py", "exec"), env)
    ns.MatchingPursuit = env["MatchingPursuit"]
    ns._saved_modules = saved
    return ns
