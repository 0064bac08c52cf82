"""ctypes binding of ``libmpb200.so`` (the C ABI declared in ``include/mpb200.h``).

There is no CPU path: if the library has not been built, or no CUDA device is
present, every entry point of this package raises.  Build with
``python -c "import __graft_entry__ as g; g.build()"`` (nvcc, sm_100a).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# MPB200_LIBRARY points the binding at an alternative build of the same sources (A/B timing of kernel variants)
LIB_PATH = os.environ.get("MPB200_LIBRARY") or os.path.join(HERE, "libmpb200.so")
SOURCES = ["mpb200.cu", "fftconv.cu", "gemm_corr.cu"]
HEADERS = ["kernels.cuh", "fused_loop.cuh", "fft_core.cuh", "bigfft.cuh", "types.h", "plan.h"]

ABI_VERSION = 130    # MPB200_VERSION of include/mpb200.h this binding was written against
MODE_AUTO, MODE_RECORRELATE, MODE_GRAM, MODE_FULL, MODE_SGRAM = 0, 1, 2, 3, 4
MODES = {"auto": MODE_AUTO, "recorrelate": MODE_RECORRELATE, "gram": MODE_GRAM, "full": MODE_FULL,
         "sgram": MODE_SGRAM}
MODE_NAMES = {v: k for k, v in MODES.items()}
# MPB200_OPT_* of include/mpb200.h (tests/test_cabi.py checks the numbers against the header)
OPT_REFRESH_EVERY, OPT_FORCE_TABLES, OPT_MAX_STEPS, OPT_POSITION_FREE, OPT_LOCAL_CONTRAST_NORM, OPT_FUSED_LOOP = 1, 2, 3, 4, 5, 6


class PlanInfo(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "n_atoms", "atom_size", "n_samples", "max_batch", "mode", "fft_size", "block", "n_blocks",
        "atom_lo", "atom_hi", "resident_batch", "fft_size2")] + [("device_bytes", C.c_uint64), ("gram_bytes", C.c_uint64)]


class MpbError(RuntimeError):
    pass


def nvcc_command(out_path: str = LIB_PATH, extra=()):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = tuple(extra) + tuple(os.environ.get("MPB_NVCC_FLAGS", "").split())
    # --cudart shared: the process (torch) already carries a CUDA runtime; linking a second, static copy into the
    # library would also drag every runtime entry point's name into the shipped artefact
    return [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--cudart", "shared",
            "-Xcompiler", "-fPIC", "-shared", *extra, "-o", out_path] + [os.path.join(CSRC, s) for s in SOURCES]


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.join(os.path.dirname(HERE), "include", "mpb200.h")]
    deps += [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))]   # any header added later
    return any(os.path.getmtime(d) > built for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the library in-tree for sm_100a (cross-compiles without a GPU)."""
    if force or needs_build():
        cmd = nvcc_command()
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        subprocess.run(cmd, check=True)
    return LIB_PATH


_lib = None

_p = C.c_void_p
_i = C.c_int
_SIGNATURES = {
    "mpb200_version": (C.c_int, []),
    "mpb200_last_error": (C.c_char_p, []),
    "mpb200_launch_count": (C.c_ulonglong, []),
    "mpb200_plan_create": (_i, [C.POINTER(_p), _i, _i, _i, _i, _i, _i, _i, C.c_uint64]),
    "mpb200_plan_destroy": (_i, [_p]),
    "mpb200_plan_info_get": (_i, [_p, C.POINTER(PlanInfo)]),
    "mpb200_plan_set_option": (_i, [_p, _i, C.c_longlong]),
    "mpb200_plan_timing_enable": (_i, [_p, _i]),
    "mpb200_plan_timing_read": (_i, [_p, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "mpb200_plan_set_dictionary": (_i, [_p, _p, _p]),
    "mpb200_plan_set_dictionary_raw": (_i, [_p, _p, _p]),
    "mpb200_plan_get_unit_dictionary": (_i, [_p, _p, _p]),
    "mpb200_sparse_code": (_i, [_p, _p, _i, _i, _p, _p, _p, _p, _p]),
    "mpb200_sparse_code_host": (_i, [_p, _p, _i, _i, _p, _p, _p, _p, _p]),
    "mpb200_correlate": (_i, [_p, _p, _i, _p, _p]),
    "mpb200_begin": (_i, [_p, _p, _i, _p]),
    "mpb200_local_best": (_i, [_p, _p, _p]),
    "mpb200_apply": (_i, [_p, _p, _p]),
    "mpb200_residual": (_i, [_p, _p, _p]),
    "mpb200_reduce_best": (_i, [_p, _i, _i, _p, _p]),
    "mpb200_scatter_add": (_i, [_p, _i, _i, _p, _i, _i, _p, _p, _p, _p, _p, _i, _p]),
    "mpb200_scatter_rows": (_i, [_p, _i, _i, _p, _i, _p, _p, _p, _i, _p]),
    "mpb200_select_dense": (_i, [_p, _i, _i, _i, _i, _p, _p]),
    "mpb200_select_lcn": (_i, [_p, _i, _i, _i, _i, _p, _p]),
    "mpb200_subtract": (_i, [_p, _i, _i, _p, _i, _i, _p, _p]),
    "mpb200_gather_atoms": (_i, [_p, _p, _i, _i, _p, _p, _i, _p]),
    "mpb200_unit_norm": (_i, [_p, _p, _i, _i, C.c_float, _p]),
    "mpb200_fold_parts": (_i, [_p, _i, _i, _i, _i, _i, _p, _p]),
    "mpb200_correlate_gemm": (_i, [_p, _i, _i, _p, _i, _i, _p, _i, _p]),
    "mpb200_dictionary_update": (_i, [_p, _i, _i, _p, _i, _i, _p, _p, _i, _p, _p, _p, _i, _p]),
    "mpb200_fft_convolve": (_i, [_p, _p, _p, _i, _i, _i, _i, C.c_float, _p, _p]),
    "mpb200_exchange_create": (_i, [_p, _i, _i, _p]),
    "mpb200_exchange_connect": (_i, [_p, _p]),
    "mpb200_exchange_mailbox": (_i, [_p, _p]),
    "mpb200_exchange_connect_local": (_i, [_p, _p]),
    "mpb200_exchange_status": (_i, [_p, _p]),
    "mpb200_exchange_disconnect": (_i, [_p]),
    "mpb200_band_limit": (_i, [_p, _i, _i, _i, _i, _i, _i, _p, _p]),
    "mpb200_spectral_band": (_i, [_p, _i, _i, _p, _i, _i, _i, _p]),
}
EXPORTED = tuple(_SIGNATURES)


def lib():
    """The loaded library; raises (never falls back) when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MpbError(
                f"{LIB_PATH} is missing: the CUDA library has not been built and this package has no "
                "CPU path.  Run: python -c \"import __graft_entry__ as g; g.build()\"")
        handle = C.CDLL(LIB_PATH)
        handle.mpb200_version.restype = C.c_int
        have = handle.mpb200_version() if hasattr(handle, "mpb200_version") else 0
        missing = [name for name in _SIGNATURES if not hasattr(handle, name)]
        if have < ABI_VERSION or missing:
            raise MpbError(
                f"{LIB_PATH} implements C ABI {have}, this package needs {ABI_VERSION}"
                + (f" (missing: {', '.join(missing)})" if missing else "")
                + "; rebuild it: python -c \"import __graft_entry__ as g; g.build()\"")
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().mpb200_last_error().decode("utf-8", "replace")
        raise MpbError(f"{what or 'mpb200 call'} failed ({rc}): {msg}")
