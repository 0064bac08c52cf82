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
    lib.emu_fft_local.argtypes = [C.c_int, C.c_int, dp, dp, dp, dp, dp, dp]
    lib.emu_fft_local.restype = C.c_int
    lib.emu_max_conflict_local.restype = C.c_int
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


def test_synthesised_gram_row_identity():
    """SGRAM never stores the K^2(2A-1) cross-correlation table: for a winner atom k* and an atom pair (d0, d1),
        G[k*, d0, l] + i G[k*, d1, l] = IFFT_M2( FFT([0^(A-1), d_k*]) * E )[l + A - 1],   |l| < A,   M2 >= 2A,
    with E the pair spectrum of the identity above at length M2, and subtracting v * d_k* at position p changes
    the correlation map of atom d by exactly -v * G[k*, d, t - p] wherever the atom fits inside the signal."""
    rng = np.random.default_rng(1)
    a, m2, n = 48, 128, 400
    dk, d0, d1 = (rng.standard_normal(a) for _ in range(3))
    s = np.fft.fft(np.concatenate([np.zeros(a - 1), dk, np.zeros(m2 - 2 * a + 1)]))
    z = np.zeros(m2, dtype=complex)
    z[:a] = d0 + 1j * d1
    y = np.fft.ifft(s * np.fft.ifft(z)) * m2

    def gram(d, lag):                                   # sum_i d_k*[i + lag] d[i]
        return sum(dk[i + lag] * d[i] for i in range(a) if 0 <= i + lag < a)

    lags = range(-(a - 1), a)
    np.testing.assert_allclose([y.real[l + a - 1] for l in lags], [gram(d0, l) for l in lags], atol=1e-9)
    np.testing.assert_allclose([y.imag[l + a - 1] for l in lags], [gram(d1, l) for l in lags], atol=1e-9)
    # the map update it stands for
    r = rng.standard_normal(n)
    p, v = 123, 0.7

    def corr(sig, d):
        pad = np.concatenate([sig, np.zeros(a)])
        return np.array([np.dot(pad[t:t + a], d) for t in range(n)])

    r2 = r.copy()
    r2[p:p + a] -= v * dk
    delta = corr(r2, d0) - corr(r, d0)
    want = np.zeros(n)
    for l in lags:
        if 0 <= p + l < n:
            want[p + l] = -v * gram(d0, l)              # map[t] = sum_i r[t+i] d[i], and r changed by -v d_k*[t+i-p]
    np.testing.assert_allclose(delta, want, atol=1e-9)


@pytest.mark.parametrize("staged", [0, 1])
@pytest.mark.parametrize("direction", [-1, 1])
def test_local_first_exchange_form(emu, direction, staged):
    """BlockFft<4096>::pass*_local: the exchange between pass 1 and pass 2 must stay inside a half-warp (the
    emulation finishes pass 1 AND pass 2 of one half-warp before it touches the next), inputs are dealt in the
    permuted table order, a staged spectrum is read from the slots pass 1 overwrites, and the result is the plain
    DFT in natural order; every shared-memory instruction of the form is bank-conflict free."""
    m = 4096
    rng = np.random.default_rng(7 + direction + staged)
    x = rng.standard_normal(m) + 1j * rng.standard_normal(m)
    want = np.fft.fft(x) if direction < 0 else np.fft.ifft(x) * m
    re, im = np.ascontiguousarray(x.real), np.ascontiguousarray(x.imag)
    ore, oim, tre, tim = np.empty(m), np.empty(m), np.empty(m), np.empty(m)
    dp = C.POINTER(C.c_double)
    rc = emu.emu_fft_local(direction, staged, re.ctypes.data_as(dp), im.ctypes.data_as(dp), ore.ctypes.data_as(dp),
                           oim.ctypes.data_as(dp), tre.ctypes.data_as(dp), tim.ctypes.data_as(dp))
    assert rc == 0
    got = ore + 1j * oim
    assert np.linalg.norm(got - want) <= 5e-7 * np.linalg.norm(want)
    # the permuted table is a permutation of the input (nibble swap inside each block of 256 bins)
    assert np.array_equal(np.sort(tre), np.sort(re))
    assert emu.emu_max_conflict_local() == 1
