"""CPU walk-through of the block FFT used by the CUDA kernels.

``csrc/fft_core.cuh`` is ``__host__ __device__``; ``csrc/host_emulation.cpp``
visits the "threads" of a block one pass at a time, so the register /
shared-memory index arithmetic the kernels rely on is checked against
``numpy.fft`` here, without a GPU.  (This is a test of product code on the
host; it is not a CPU path of the product.)"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "matching-pursuit_b200", "csrc")
SIZES = [256, 512, 1024, 2048, 4096, 8192]


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("emu") / "libfftemu.so")
    subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-I", CSRC, "-o", out,
                    os.path.join(CSRC, "host_emulation.cpp")], check=True)
    lib = C.CDLL(out)
    dp = C.POINTER(C.c_double)
    lib.emu_fft.argtypes = [C.c_int, C.c_int, C.c_int, dp, dp, dp, dp]
    lib.emu_fft.restype = C.c_int
    lib.emu_max_conflict.argtypes = [C.c_int, C.c_int]
    lib.emu_max_conflict.restype = C.c_int
    return lib


def run(lib, x, direction, use_double):
    m = x.shape[0]
    re, im = np.ascontiguousarray(x.real, dtype=np.float64), np.ascontiguousarray(x.imag, dtype=np.float64)
    ore, oim = np.empty(m), np.empty(m)
    dp = C.POINTER(C.c_double)
    rc = lib.emu_fft(m, direction, int(use_double), re.ctypes.data_as(dp), im.ctypes.data_as(dp),
                     ore.ctypes.data_as(dp), oim.ctypes.data_as(dp))
    assert rc == 0
    return ore + 1j * oim


@pytest.mark.parametrize("m", SIZES)
@pytest.mark.parametrize("direction", [-1, 1])
def test_block_fft_matches_numpy(emu, m, direction):
    rng = np.random.default_rng(m + direction)
    x = rng.standard_normal(m) + 1j * rng.standard_normal(m)
    want = np.fft.fft(x) if direction < 0 else np.fft.ifft(x) * m
    got64 = run(emu, x, direction, True)
    assert np.abs(got64 - want).max() <= 1e-10 * np.abs(want).max()
    got32 = run(emu, x, direction, False)
    # fp32 butterflies: relative L2 error of a few ulp * sqrt(log2 M)
    assert np.linalg.norm(got32 - want) <= 5e-7 * np.linalg.norm(want)


@pytest.mark.parametrize("m", SIZES)
def test_shared_layout_is_conflict_free(emu, m):
    # worst number of 8-byte words of a half-warp that fall in one bank, per pass
    for p in (1, 2, 3):
        assert emu.emu_max_conflict(m, p) <= 2, (m, p)


def test_pair_packing_identity():
    """The kernels correlate a real window with TWO atoms per complex inverse
    transform: with E = FFT_inverse_kernel(d0 + i*d1)/M,  IFFT(X * E) has the
    correlation with d0 in its real part and with d1 in its imaginary part."""
    rng = np.random.default_rng(0)
    m, a = 512, 100
    x = np.zeros(m)
    x[:m] = rng.standard_normal(m)
    d0, d1 = rng.standard_normal(a), rng.standard_normal(a)
    z = np.zeros(m, dtype=complex)
    z[:a] = d0 + 1j * d1
    e = np.fft.ifft(z)                      # = sum_j z[j] exp(+2 pi i j m / M) / M
    y = np.fft.ifft(np.fft.fft(x) * e) * m  # unnormalised inverse, as the kernels do
    valid = m - a + 1
    want0 = np.array([np.dot(x[t:t + a], d0) for t in range(valid)])
    want1 = np.array([np.dot(x[t:t + a], d1) for t in range(valid)])
    np.testing.assert_allclose(y.real[:valid], want0, atol=1e-9)
    np.testing.assert_allclose(y.imag[:valid], want1, atol=1e-9)
