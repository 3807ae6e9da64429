"""Host-side mirror of the reference's sort worker (src/worker/SortWorker.js): same factory, same message protocol.

    worker = createSortWorker(splatCount, useSharedMemory, enableSIMDInSort, integerBasedSort, dynamicMode,
                              splatSortDistanceMapPrecision)                       # SortWorker.js:202-256
    worker.onmessage = lambda e: ...                                               # Viewer.js:1243-1298
    worker.postMessage({'centers': ..., 'sceneIndexes': ..., 'range': {...}})      # SortWorker.js:84-98
    worker.postMessage({'sort': {...}})                                            # SortWorker.js:99-113
    worker.terminate()                                                             # Viewer.js:1311

Messages are plain dicts with the reference's keys; `e.data` carries the payload like a MessageEvent.  The wasm
module + WebAssembly.Memory of the reference are replaced by one gs_engine on a B200; the "shared memory" views the
worker hands back in `sortSetupPhase1Complete` are page-locked host buffers the engine copies from / into.
No arithmetic happens here: distances, buckets and the scatter all run in libgsplat_b200.so's CUDA kernels.
"""
from __future__ import annotations

import time
from types import SimpleNamespace

import numpy as np

from . import _native as N
from .engine import Engine

DefaultSplatSortDistanceMapPrecision = 16  # src/Constants.js:3
MaxScenes = 32                             # src/Constants.js:7


class SortWorker:
    def __init__(self, device: int = 0):
        self.onmessage = None
        self._device = device
        self._engine: Engine | None = None
        self._splat_count = 0
        self._use_shared = False
        self._integer = True
        self._dynamic = False
        self._range = 1 << DefaultSplatSortDistanceMapPrecision
        self._uploaded = 0
        self._shared = {}

    # -- Worker API --------------------------------------------------------------------------------------------
    def postMessage(self, data: dict) -> None:  # noqa: N802 (reference spelling)
        if "centers" in data and data["centers"] is not None:
            self._on_centers(data)
        elif "sort" in data and data["sort"]:
            self._on_sort(data["sort"])
        elif "init" in data and data["init"]:
            self._on_init(data["init"])

    def terminate(self) -> None:
        if self._engine is not None:
            self._engine.close()
            self._engine = None

    def _emit(self, payload: dict) -> None:
        if self.onmessage is not None:
            self.onmessage(SimpleNamespace(data=payload))

    # -- handlers ----------------------------------------------------------------------------------------------
    def _on_init(self, init: dict) -> None:  # SortWorker.js:114-197
        self._splat_count = int(init["splatCount"])
        self._use_shared = bool(init.get("useSharedMemory", False))
        self._integer = bool(init.get("integerBasedSort", True))
        self._dynamic = bool(init.get("dynamicMode", False))
        self._range = int(init.get("distanceMapRange", 1 << DefaultSplatSortDistanceMapPrecision))
        self._uploaded = 0
        self._engine = Engine(self._splat_count, device=self._device, distance_map_range=self._range,
                              integer_based_sort=self._integer, dynamic_mode=self._dynamic)
        msg = {"sortSetupPhase1Complete": True}
        if self._use_shared:  # SortWorker.js:180-191: views the main thread writes / reads directly
            n = max(self._splat_count, 1)
            self._shared = {
                "indexesToSort": N.pinned_empty(n, np.uint32),
                "sortedIndexes": N.pinned_empty(n, np.uint32),
                "precomputedDistances": N.pinned_empty(n, np.int32 if self._integer else np.float32),
                "transforms": N.pinned_empty(16 * MaxScenes, np.float32),
            }
            msg.update({
                "indexesToSortBuffer": self._shared["indexesToSort"], "indexesToSortOffset": 0,
                "sortedIndexesBuffer": self._shared["sortedIndexes"], "sortedIndexesOffset": 0,
                "precomputedDistancesBuffer": self._shared["precomputedDistances"], "precomputedDistancesOffset": 0,
                "transformsBuffer": self._shared["transforms"], "transformsOffset": 0,
            })
        self._emit(msg)

    def _on_centers(self, data: dict) -> None:  # SortWorker.js:84-98
        rng = data["range"]
        start, count = int(rng["from"]), int(rng["count"])
        want = np.int32 if self._integer else np.float32
        centers = np.frombuffer(data["centers"], dtype=want) if not isinstance(data["centers"], np.ndarray) else data["centers"].view(want)
        scene = data.get("sceneIndexes") if self._dynamic else None
        if scene is not None and not isinstance(scene, np.ndarray):
            scene = np.frombuffer(scene, dtype=np.uint32)
        self._engine.upload_centers(centers.reshape(-1)[: count * 4], None if scene is None else scene[:count], start)
        self._uploaded = start + count

    def _on_sort(self, s: dict) -> None:  # SortWorker.js:31-81, 99-113
        t0 = time.perf_counter()
        render_count = min(int(s.get("splatRenderCount") or 0), self._uploaded)
        sort_count = min(int(s.get("splatSortCount") or 0), self._uploaded)
        use_pre = bool(s.get("usePrecomputedDistances", False))
        if self._use_shared:
            indexes, transforms = self._shared["indexesToSort"], self._shared["transforms"]
            pre = self._shared["precomputedDistances"] if use_pre else None
            out = self._shared["sortedIndexes"]
        else:
            indexes, transforms = s.get("indexesToSort"), s.get("transforms")
            pre = s.get("precomputedDistances") if use_pre else None
            out = np.empty(max(render_count, 1), np.uint32)
        sorted_idx, _ = self._engine.sort(np.asarray(s["modelViewProj"], np.float32), sort_count, render_count,
                                          None if indexes is None else indexes[:render_count],
                                          transforms=transforms if self._dynamic else None, precomputed=pre, out=out[: max(render_count, 0)])
        msg = {"sortDone": True, "splatSortCount": sort_count, "splatRenderCount": render_count, "sortTime": 0}
        if not self._use_shared:
            msg["sortedIndexes"] = sorted_idx
        msg["sortTime"] = (time.perf_counter() - t0) * 1000.0
        self._emit(msg)


def createSortWorker(splatCount, useSharedMemory, enableSIMDInSort, integerBasedSort, dynamicMode,  # noqa: N802,N803
                     splatSortDistanceMapPrecision=DefaultSplatSortDistanceMapPrecision, device: int = 0) -> SortWorker:
    """SortWorker.js:202-256.  `enableSIMDInSort` selects between two spellings of the same arithmetic in the reference
    (sorter.cpp vs sorter_no_simd.cpp); there is one CUDA path here, so it is accepted and ignored."""
    del enableSIMDInSort
    worker = SortWorker(device=device)
    worker._pending_init = {  # posted immediately like the reference does (SortWorker.js:237-254)
        "init": {"splatCount": splatCount, "useSharedMemory": useSharedMemory, "integerBasedSort": integerBasedSort,
                 "dynamicMode": dynamicMode, "distanceMapRange": 1 << splatSortDistanceMapPrecision,
                 "Constants": {"BytesPerFloat": 4, "BytesPerInt": 4, "MemoryPageSize": 65536, "MaxScenes": MaxScenes}}}
    return worker


def start(worker: SortWorker) -> None:
    """Deliver the queued 'init' (call after assigning worker.onmessage; a JS Worker does this on the next tick)."""
    init = getattr(worker, "_pending_init", None)
    if init is not None:
        worker._pending_init = None
        worker.postMessage(init)
