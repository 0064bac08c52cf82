"""Host-side logic that needs no GPU: shard arithmetic as properties, the plan cache's eviction policy (with
stand-in plans), the slice -> bin-progression mapping of the band limiter, and the E=32 variant of the FFT core's
index maps."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch
from hypothesis import given, settings, strategies as st

import matching_pursuit_b200 as mpb
from matching_pursuit_b200 import matchingpursuit as mmp
from matching_pursuit_b200.distributed import atom_range, shard_batch


@settings(max_examples=200, deadline=None)
@given(st.integers(1, 5000), st.integers(1, 64))
def test_shards_partition_the_range(total, world):
    """Contiguous, ordered, balanced (sizes differ by at most one) and complete."""
    parts = [shard_batch(total, world, r) for r in range(world)]
    assert parts[0][0] == 0 and parts[-1][1] == total
    assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
    sizes = [hi - lo for lo, hi in parts]
    assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    if world <= total:
        assert [atom_range(total, world, r) for r in range(world)] == parts
    else:
        with pytest.raises(ValueError):
            atom_range(total, world, 0)


class _FakePlan:
    closed = 0

    def __init__(self, nbytes, max_batch=8, pins=0):
        self.device_bytes, self.max_batch, self.pins = nbytes, max_batch, pins

    def close(self):
        _FakePlan.closed += 1


def test_plan_cache_evicts_oldest_first_within_budget():
    saved = dict(mmp._PLAN_CACHE)
    mmp._PLAN_CACHE.clear()
    try:
        _FakePlan.closed = 0
        for i in range(5):
            mmp._PLAN_CACHE[(0, i)] = _FakePlan(100)
        mmp._PLAN_CACHE[(1, 0)] = _FakePlan(10_000)                  # another device: not counted against device 0
        mmp._evict_until(0, max_bytes=350, max_entries=10)
        # device 0's oldest idle entries go first until it is within its budget; other devices are left alone
        assert [k for k in mmp._PLAN_CACHE if k[0] == 0] == [(0, 2), (0, 3), (0, 4)]
        assert _FakePlan.closed == 2 and (1, 0) in mmp._PLAN_CACHE
        mmp._evict_until(0, max_bytes=10 ** 9, max_entries=2)
        assert [k for k in mmp._PLAN_CACHE if k[0] == 0] == [(0, 3), (0, 4)] and (1, 0) in mmp._PLAN_CACHE
        # a plan somebody holds (`with plan:`) is never closed, however far over budget the cache is
        mmp._PLAN_CACHE[(0, 3)].pins = 1
        mmp._evict_until(0, max_bytes=0, max_entries=0)
        assert [k for k in mmp._PLAN_CACHE if k[0] == 0] == [(0, 3)] and (1, 0) in mmp._PLAN_CACHE
        mmp.clear_plan_cache(0)
        assert (0, 3) in mmp._PLAN_CACHE and (1, 0) in mmp._PLAN_CACHE
        mmp._PLAN_CACHE[(0, 3)].pins = 0
        mmp.clear_plan_cache()
        assert not mmp._PLAN_CACHE
    finally:
        mmp._PLAN_CACHE.clear()
        mmp._PLAN_CACHE.update(saved)


@settings(max_examples=300, deadline=None)
@given(st.integers(2, 400).map(lambda h: 2 * h), st.one_of(st.none(), st.integers(-500, 500)),
       st.one_of(st.none(), st.integers(-500, 500)), st.one_of(st.none(), st.integers(1, 9)))
def test_slice_maps_to_an_arithmetic_bin_progression(length, start, stop, step):
    """engine.band_limit hands (first bin, step, count) to the library: it must select exactly the bins the
    reference's tensor slice selects (modules/conv.py:24-29)."""
    slce = slice(start, stop, step)
    n_bins = length // 2 + 1
    bins = range(*slce.indices(n_bins))
    want = np.arange(n_bins)[slce]
    assert list(bins) == want.tolist()
    if len(bins):
        assert bins[0] >= 0 and bins[0] + (len(bins) - 1) * bins.step <= length // 2


def test_fft_core_e32_index_maps(tmp_path):
    """BlockFft<4096, float, 32> (the 128-thread variant of k_delta's transform): input and output index maps
    must each be a permutation of 0..M-1 -- compiled for the host from the same header the kernels use."""
    src = tmp_path / "e32.cpp"
    src.write_text('''
#include <cstdio>
#include <vector>
#include "fft_core.cuh"
using F = mpb::BlockFft<4096, float, 32>;
int main() {
    static_assert(F::E == 32 && F::T == 128 && F::NB1 == 2 && F::NB2 == 2, "geometry");
    std::vector<int> in(F::M, 0), out(F::M, 0);
    for (int t = 0; t < F::T; ++t)
        for (int e = 0; e < F::E; ++e) { in[F::in_index(t, e)]++; out[F::out_index(t, e)]++; }
    for (int i = 0; i < F::M; ++i) if (in[i] != 1 || out[i] != 1) { std::printf("bad %d\\n", i); return 1; }
    std::printf("ok\\n");
    return 0;
}
''')
    exe = tmp_path / "e32"
    csrc = os.path.join(os.path.dirname(os.path.abspath(mpb.__file__)), "csrc")
    subprocess.run(["g++", "-std=c++17", "-O1", "-I", csrc, str(src), "-o", str(exe)], check=True)
    assert subprocess.run([str(exe)], capture_output=True, text=True).stdout.strip() == "ok"


def test_event_lists_materialise_on_first_access():
    """Event lists carry packed arrays and build the reference's tuples (modules/matchingpursuit.py:305-321) only
    when somebody looks at them; until then len() answers from the arrays, afterwards they are ordinary lists."""
    import copy
    import pickle
    from matching_pursuit_b200.matchingpursuit import EventList, flatten_atom_dict
    e, a = 5, 4

    def make():
        return EventList.from_packed(torch.arange(e), torch.zeros(e, dtype=torch.int64), torch.arange(e) * 10,
                                     torch.randn(e, a))

    ev = make()
    assert isinstance(ev, list) and len(ev) == 5 and ev._lazy and bool(ev)
    item = ev[2]
    assert not ev._lazy and item[0] == 2 and item[1] == 0 and int(item[2]) == 20
    assert item[2].shape == (1, 1) and item[2].dtype == torch.int64 and item[3].shape == (1, 1, a)
    assert [x[0] for x in make()] == [0, 1, 2, 3, 4]
    grown = []
    grown.extend(make())
    assert len(grown) == 5 and len(make()[1:3]) == 2 and len(make() + [1]) == 6 and len(list(make())) == 5
    assert len(pickle.loads(pickle.dumps(make()))) == 5 and len(copy.copy(make())) == 5
    assert len(flatten_atom_dict({1: make(), 2: make()})) == 10
    assert len(sorted(make(), key=lambda x: -x[0])) == 5
    mutated = make()
    mutated.append("x")
    assert len(mutated) == 6 and mutated[-1] == "x"
    z = torch.zeros(0, dtype=torch.int64)
    empty = EventList.from_packed(z, z, z, torch.zeros(0, a))
    assert len(empty) == 0 and not empty and list(empty) == []
