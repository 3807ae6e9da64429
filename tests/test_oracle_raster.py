"""CPU: the raster restatement (oracle/raster_oracle.c) and the host packing/uniform code still produce the committed fixture
(tests/golden/raster_golden.npz, written by tests/golden/make_raster_golden.py)."""
import sys
from pathlib import Path

import numpy as np
import pytest

sys.path.insert(0, str(Path(__file__).resolve().parent / "golden"))
import raster_cases  # noqa: E402

GOLD = np.load(Path(__file__).resolve().parent / "golden" / "raster_golden.npz")


@pytest.mark.parametrize("name", list(raster_cases.CASES))
def test_raster_oracle_reproduces_fixture(oracle_mod, name):
    order, frame, ps = raster_cases.oracle_outputs(name, oracle_mod)
    assert np.array_equal(order, GOLD[name + "|order"])
    assert np.array_equal(np.packbits(ps["valid"].astype(np.uint8)), GOLD[name + "|valid"])
    want = GOLD[name + "|frame"]
    assert frame.shape == want.shape
    # same compiler flags -> identical; leave room for another libm's expf
    assert np.abs(frame - want).max() < 2e-6
    assert want[..., 3].max() > 0.9 and (want[..., 3] == 0).any()


def test_fixture_order_is_a_permutation_back_to_front(oracle_mod):
    """Sanity of the stored draw order: a permutation along which the reference's integer distance (sorter.cpp:64-74) falls from
    bucket to bucket and, inside a bucket, later input positions come first (sorter.cpp:158-167)."""
    del oracle_mod
    for name in raster_cases.CASES:
        v, raw = raster_cases.host_viewer(name)
        order = GOLD[name + "|order"].astype(np.int64)
        assert np.array_equal(np.sort(order), np.arange(raw.count))
        m = v.mvp_matrix().astype(np.float32)
        row = np.array([int(np.float64(m[2]) * 1000.0), int(np.float64(m[6]) * 1000.0), int(np.float64(m[10]) * 1000.0)], np.int64)
        d = (v.splatMesh.packed.int_centers[:, :3].astype(np.int64) * row[None, :]).sum(axis=1)
        assert np.abs(d).max() < 2**31
        rm = np.float32(65535) / (np.float32(d.max()) - np.float32(d.min()))
        b = ((d - d.min()).astype(np.float32) * rm).astype(np.int64)[order]
        assert (np.diff(b) <= 0).all()
        same = np.diff(b) == 0
        assert (np.diff(order)[same] < 0).all()
