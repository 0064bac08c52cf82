"""CPU-side checks of the drop-in boundary: the C-ABI library builds for
sm_100a, loads, and exports every symbol ``include/mpb200.h`` declares; the
product refuses to run without a CUDA device (no CPU path)."""
import ctypes
import os
import re

import pytest
import torch

import matching_pursuit_b200 as mpb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mpb200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mpb200_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def library():
    mpb.build()
    return ctypes.CDLL(mpb.LIB_PATH)


def test_header_declares_the_boundary():
    syms = declared_symbols()
    for required in ("mpb200_plan_create", "mpb200_plan_set_dictionary", "mpb200_sparse_code",
                     "mpb200_sparse_code_host", "mpb200_correlate", "mpb200_begin", "mpb200_local_best",
                     "mpb200_apply", "mpb200_reduce_best", "mpb200_scatter_add", "mpb200_fft_convolve"):
        assert required in syms


def test_library_exports_every_declared_symbol(library):
    for name in declared_symbols():
        assert hasattr(library, name), f"{name} is declared in include/mpb200.h but not exported"


def test_binding_covers_every_declared_symbol():
    assert sorted(mpb.EXPORTED) == declared_symbols()


def test_version(library):
    """Header, library and binding agree on the ABI version; no closed batch-copy entry point is named anywhere in
    the shipped artefact (the CUDA runtime is linked shared, not static)."""
    library.mpb200_version.restype = ctypes.c_int
    header = int(re.search(r"#define\s+MPB200_VERSION\s+(\d+)", open(HEADER).read()).group(1))
    from importlib import import_module
    assert library.mpb200_version() == header == import_module("matching_pursuit_b200._lib").ABI_VERSION
    blob = open(mpb.LIB_PATH, "rb").read()
    assert b"MemcpyBatchAsync" not in blob


def test_signatures_carry_no_torch_types():
    text = open(HEADER).read()
    assert "torch" not in re.sub(r"/\*.*?\*/", "", text, flags=re.S).lower()
    assert "at::" not in text and "c10::" not in text


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    with pytest.raises(mpb.MpbError):
        mpb.Plan(4, 16, 128, 1)
    sig = torch.zeros(1, 1, 128)
    d = torch.randn(4, 16)
    with pytest.raises(mpb.MpbError):
        mpb.sparse_code(sig, d, n_steps=2)


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "matching-pursuit_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "/root/reference" not in src, f


def test_first_seen_grouping_matches_reference_order():
    import numpy as np
    from importlib import import_module
    mp_mod = import_module("matching_pursuit_b200.matchingpursuit")
    atoms = np.array([5, 2, 5, 7, 2, 2, 9, 5])
    perm, seen = mp_mod._first_seen_grouping(atoms)
    assert seen.tolist() == [5, 2, 7, 9]
    assert perm.tolist() == [0, 2, 7, 1, 4, 5, 3, 6]


def test_option_and_mode_numbers_match_the_header():
    """The binding's MPB200_OPT_* / MPB200_MODE_* numbers are the header's."""
    from matching_pursuit_b200 import _lib
    text = open(HEADER).read()
    opts = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define\s+MPB200_OPT_([A-Z_]+)\s+(\d+)", text)}
    assert opts, "no MPB200_OPT_* in the header"
    for name, value in opts.items():
        assert getattr(_lib, "OPT_" + name) == value, name
    modes = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define\s+MPB200_MODE_([A-Z_]+)\s+(\d+)", text)}
    for name, value in modes.items():
        assert getattr(_lib, "MODE_" + name) == value, name
    version = int(re.search(r"#define\s+MPB200_VERSION\s+(\d+)", text).group(1))
    assert _lib.ABI_VERSION == version
