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
    ap.add_argument("--auto-split", action="store_true", help="torchrun mode: keep the engine's split threshold (default: always split, GS_SHARD_MIN=0)")
    a = ap.parse_args()
    import os
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        if not a.auto_split:
            os.environ["GS_SHARD_MIN"] = "0"
        return main_sharded(a)
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


def main_sharded(a):
    """Under torchrun: ONE sort split over the ranks by input position (gs_sort_sharded); result assembled in rank 0's memory.
    Time = CUDA events on every rank's stream around its part, max over ranks (rank 0's includes the wait for all peers)."""
    import ctypes as C
    import os
    import torch
    import torch.distributed as dist
    import cases
    import gaussiansplats3d_b200 as gs
    from gaussiansplats3d_b200.parallel import ShardedSort
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    for m in [int(x) for x in a.sizes.split(",")]:
        n = m * 1_000_000
        c = cases.sort_case(seed=10 + m, n=n, index_kind="shuffled" if a.shuffled else "identity")
        with gs.Engine(n, device=local, rank=rank, world_size=world) as e:
            e.upload_centers(c["centers"])
            idx_dev = None
            want, _ = e.sort(c["mvp"], n, n, c["indexes"] if a.shuffled else None)      # single-GPU answer (also uploads the index list)
            if a.shuffled:
                idx_dev, _ = e.buffer_dev(gs._native.GS_BUF_INDEXES_TO_SORT)
            ss = ShardedSort(e, rank, world)
            p = e._sort_params(c["mvp"], n, n, None, None, None, indexes_dev=idx_dev)
            out = np.empty(n, np.uint32) if rank == 0 else None
            gs._native.check(e._lib.gs_sort_sharded(e._h, C.byref(p), gs._native.ptr(out), None), "gs_sort_sharded")
            exact = bool(np.array_equal(out, want)) if rank == 0 else None
            ev0, ev1 = e.event(), e.event()
            times = []
            for _ in range(a.reps + 2):
                e.flush_l2()
                e.synchronize()
                dist.barrier()
                ev0.record()
                gs._native.check(e._lib.gs_sort_sharded(e._h, C.byref(p), None, None), "gs_sort_sharded")
                ev1.record()
                e.synchronize()
                t = torch.tensor([ev0.elapsed_ms(ev1)], device="cuda", dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                times.append(float(t.item()))
            dist.barrier()
        if rank == 0:
            ms = float(np.median(times[2:]))
            print(json.dumps({"splats": n, "gpus": world, "indexes": "shuffled" if a.shuffled else "identity", "sort_ms": ms, "msplats_per_s": n / ms / 1e3,
                              "bit_exact_vs_single_gpu": exact, "split": "engine threshold" if a.auto_split else "forced", "exchange": "peer memory (min/max, run lengths, 4 B/splat into rank 0)"}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
