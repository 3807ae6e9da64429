"""Generates tests/golden/sort_golden.npz by running the REFERENCE's own sorter (oracle/_ref, compiled from
/root/reference/src/worker/sorter_no_simd.cpp by oracle/Makefile) on seeded inputs.  Run in the container that has
/root/reference:   python tests/golden/make_golden.py
Only the seeds/arguments and a checksum + the full output of small cases are stored; inputs are regenerated from seeds."""
import hashlib
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import cases  # noqa: E402
import oracle  # noqa: E402


def main():
    oracle.build()
    assert oracle.have_ref(), "oracle/_ref missing: needs /root/reference"
    store = {}
    names = []
    for name, kw in cases.sort_matrix(n=3000, seeds=(11,)):
        c = cases.sort_case(**kw)
        for R in cases.RANGES:
            if not c["integer_sort"] and R == (1 << 20):
                R = 1 << 22  # float mode allows up to 24 bits (Viewer.js:208-210)
            out, mapped, freq = oracle.ref_sort_indexes(*cases.call_args(c, R), want_scratch=True)
            key = f"{name}|R{R}"
            names.append(key)
            store[key + "|out"] = out.astype(np.uint32)
            store[key + "|sha"] = np.frombuffer(hashlib.sha256(out.tobytes() + mapped.tobytes()).digest(), np.uint8)
    store["names"] = np.array(names)
    np.savez_compressed(ROOT / "tests" / "golden" / "sort_golden.npz", **store)
    print(f"wrote {len(names)} cases")


if __name__ == "__main__":
    main()
