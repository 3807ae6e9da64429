"""Writes tests/golden/raster_golden.npz: draw order (sort oracle) and float frame (raster oracle) of the seeded cases in
raster_cases.py.  Run here or anywhere the oracle builds:   python tests/golden/make_raster_golden.py"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests" / "golden"))
import oracle  # noqa: E402
import raster_cases  # noqa: E402


def main():
    oracle.build()
    store = {}
    for name in raster_cases.CASES:
        order, frame, ps = raster_cases.oracle_outputs(name, oracle)
        if oracle.have_ref():   # the draw order stored is the compiled reference's own
            v, raw = raster_cases.host_viewer(name)
            n = raw.count
            ref = oracle.ref_sort_indexes(np.arange(n, dtype=np.uint32), v.splatMesh.packed.int_centers, None, v.mvp_matrix().astype(np.float32), None, None,
                                          1 << 16, n, n, n, False, True, False)
            assert np.array_equal(ref, order), name
        store[name + "|order"] = order.astype(np.uint32)
        store[name + "|frame"] = frame.astype(np.float32)
        store[name + "|valid"] = np.packbits(ps["valid"].astype(np.uint8))
        print(name, "visible", int(ps["valid"].sum()), "alpha max", float(frame[..., 3].max()))
    np.savez_compressed(ROOT / "tests" / "golden" / "raster_golden.npz", **store)


if __name__ == "__main__":
    main()
