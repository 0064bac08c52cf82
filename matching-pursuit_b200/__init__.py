"""matching-pursuit_b200 -- B200-native greedy matching pursuit behind the
reference's PyTorch entry points.

The directory name carries a hyphen (it mirrors the reference repository's
name), so import it through the root-level shim::

    import matching_pursuit_b200 as mpb
    events, scatter, residual = mpb.sparse_code(signal, d, n_steps=64, flatten=True, return_residual=True)

Only what the hot path needs lives here: ``csrc/`` (CUDA kernels + the C ABI of
``include/mpb200.h``), the ctypes binding (``_lib``), the plan wrapper
(``engine``) and host-side mirrors of the reference modules
(``matchingpursuit``, ``conv``, ``fft``, ``decompose``, ``multibanddict``,
``mp``).  There is no CPU path: without the built library or a CUDA device
every entry point raises :class:`MpbError`.
"""
from ._lib import MpbError, build, lib, LIB_PATH, EXPORTED  # noqa: F401
from .engine import Plan, reduce_best, unpack_best, unit_norm, gather_atoms, scatter_add, scatter_rows, \
    select_dense, subtract  # noqa: F401
from .matchingpursuit import (sparse_code, sparse_code_arrays, sparse_feature_map, build_scatter_segments,  # noqa: F401
                              flatten_atom_dict, dictionary_learning_step, get_plan, clear_plan_cache, EventList)

from . import autograd, conv, decompose, distributed, fft, mp, multibanddict, sparse  # noqa: F401,E402
from .multibanddict import BandSpec, MultibandDictionaryLearning  # noqa: F401,E402
from .sparse import sparsify2  # noqa: F401,E402

__version__ = "0.1.0"
