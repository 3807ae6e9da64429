#!/usr/bin/env python
"""BASELINE config 5: sort-only throughput sweep (1M..16M splats, integer mode, R = 65536, identity and shuffled indexes).
Prints one JSON line per size with the per-kernel timeline; `--cpu` adds the compiled reference sorter on one host core."""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="1,2,4,8,16")
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--cpu", action="store_true")
    ap.add_argument("--shuffled", action="store_true")
    a = ap.parse_args()
    import cases
    import gaussiansplats3d_b200 as gs
    peak = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0
    for m in [int(x) for x in a.sizes.split(",")]:
        n = m * 1_000_000
        c = cases.sort_case(seed=10 + m, n=n, index_kind="shuffled" if a.shuffled else "identity")
        with gs.Engine(n) as e:
            e.upload_centers(c["centers"])
            idx = c["indexes"] if a.shuffled else None
            idx_dev = None
            if a.shuffled:  # keep the index list resident: upload once through a first sort, then reuse the device copy
                e.sort(c["mvp"], n, n, idx, download=False)
                idx_dev, _ = e.buffer_dev(gs._native.GS_BUF_INDEXES_TO_SORT)
            for _ in range(3):
                e.sort(c["mvp"], n, n, None if idx_dev is None else None, download=False) if idx_dev is None else e._lib and None
            ev0, ev1 = e.event(), e.event()
            times = []
            for _ in range(a.reps):
                e.flush_l2()
                ev0.record()
                if idx_dev is None:
                    p = e._sort_params(c["mvp"], n, n, None, None, None)
                else:
                    p = e._sort_params(c["mvp"], n, n, None, None, None, indexes_dev=idx_dev)
                import ctypes as C
                gs._native.check(e._lib.gs_sort(e._h, C.byref(p), None, None), "gs_sort")
                ev1.record()
                times.append(ev0.elapsed_ms(ev1))
            e.set_profiling(True)
            gs._native.check(e._lib.gs_sort(e._h, C.byref(p), None, None), "gs_sort")
            kt = dict(e.kernel_timings())
            e.set_profiling(False)
        ms = float(np.median(times))
        line = {"splats": n, "indexes": "shuffled" if a.shuffled else "identity", "sort_ms": ms, "msplats_per_s": n / ms / 1e3,
                "algorithmic_gbs": 24 * n / ms / 1e6, "frac_of_hbm_peak": 24 * n / ms / 1e6 / peak, "kernel_ms": {k: round(v, 4) for k, v in kt.items()}}
        if a.cpu:
            import oracle
            t0 = time.perf_counter()
            oracle.ref_sort_indexes(*cases.call_args(c, 1 << 16))
            line["cpu_reference_ms"] = (time.perf_counter() - t0) * 1e3
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
