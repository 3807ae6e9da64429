"""GPU parity: the CUDA sort (through the C ABI) must reproduce the reference's sortIndexes() bit for bit."""
import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu


def _expect(oracle_mod, c, R):
    if oracle_mod.have_ref():
        return oracle_mod.ref_sort_indexes(*cases.call_args(c, R))
    return oracle_mod.port_sort_indexes(*cases.call_args(c, R))


@pytest.mark.parametrize("name,kw", cases.sort_matrix(n=20000, seeds=(0,)))
def test_dropin_sortIndexes_bit_exact(gs, oracle_mod, name, kw):
    c = cases.sort_case(**kw)
    for R in cases.RANGES + ((1 << 22,) if not c["integer_sort"] else ()):
        want = _expect(oracle_mod, c, R)
        got = gs.sort_indexes(*cases.call_args(c, R))
        assert np.array_equal(got, want), f"{name} R={R}: first mismatch at {np.flatnonzero(got != want)[:5]}"


def test_dropin_scratch_outputs(gs, oracle_mod):
    """mappedDistances / frequencies hold what the reference leaves there."""
    c = cases.sort_case(seed=3, n=50000, index_kind="shuffled", sort_frac=0.6)
    R = 1 << 16
    out, mapped, freq = gs.sort_indexes(*cases.call_args(c, R), want_scratch=True)
    want, buckets = oracle_mod.port_sort_indexes(*cases.call_args(c, R), want_buckets=True)
    s0 = c["render_count"] - c["sort_count"]
    assert np.array_equal(out, want)
    assert np.array_equal(mapped[s0:], buckets[s0:])
    counts = np.bincount(buckets[s0:], minlength=R)
    assert np.array_equal(freq, (np.cumsum(counts) - counts).astype(np.uint32))
    if oracle_mod.have_ref():
        _, rmapped, rfreq = oracle_mod.ref_sort_indexes(*cases.call_args(c, R), want_scratch=True)
        assert np.array_equal(mapped[s0:], rmapped[s0:]) and np.array_equal(freq, rfreq)


def test_golden_vectors(gs):
    """Committed outputs of the reference's compiled sorter (tests/golden/make_golden.py)."""
    gold = np.load(cases.__file__.replace("cases.py", "golden/sort_golden.npz"))
    table = dict(cases.sort_matrix(n=3000, seeds=(11,)))
    for key in [str(k) for k in gold["names"]]:
        name, rtag = key.split("|")
        c = cases.sort_case(**table[name])
        got = gs.sort_indexes(*cases.call_args(c, int(rtag[1:])))
        assert np.array_equal(got, gold[key + "|out"]), key


@pytest.mark.parametrize("integer,dynamic", [(True, False), (False, False), (True, True), (False, True)])
def test_engine_worker_protocol(gs, oracle_mod, integer, dynamic):
    """init -> centers (in two ranges) -> sort (partial, then full), like the Viewer drives the worker."""
    n = 150_000
    c = cases.sort_case(seed=9, n=n, integer=integer, dynamic=dynamic, index_kind="octree")
    R = 1 << 16
    with gs.Engine(n, distance_map_range=R, integer_based_sort=integer, dynamic_mode=dynamic) as e:
        half = n // 2
        si = c["scene_indexes"]
        e.upload_centers(c["centers"][:half], None if si is None else si[:half], 0)
        e.upload_centers(c["centers"][half:], None if si is None else si[half:], half)
        for sort_count in (c["render_count"] // 8, c["render_count"]):
            c2 = dict(c, sort_count=sort_count)
            out, ms = e.sort(c["mvp"], sort_count, c["render_count"], c["indexes"], transforms=c["transforms"])
            assert np.array_equal(out, _expect(oracle_mod, c2, R))
            assert ms > 0
        t = e.timings()
        assert t["kernel_launches"] >= 4


def test_identity_fast_path_and_precomputed(gs, oracle_mod):
    n = 300_000
    c = cases.sort_case(seed=21, n=n, index_kind="identity")
    with gs.Engine(n) as e:
        e.upload_centers(c["centers"])
        out, _ = e.sort(c["mvp"], n, n, None)          # NULL indexes = identity (Viewer.js:2061-2074)
        assert np.array_equal(out, _expect(oracle_mod, c, 1 << 16))
        # D1 + usePrecomputedDistances branch: integer rows are Math.round of the f64 matrix
        mvp64 = c["mvp"].astype(np.float64)
        d = e.compute_distances(mvp64, n)
        rows = np.floor(mvp64[[2, 6, 10]] * 1000.0 + 0.5).astype(np.int64)
        want_d = (c["centers"][:, :3].astype(np.int64) * rows).sum(1)
        want_d = ((want_d + 2**31) % 2**32 - 2**31).astype(np.int32)
        assert np.array_equal(d, want_d)
        out2, _ = e.sort(c["mvp"], n, n, None, precomputed=d)
        c3 = dict(c, precomputed=d, use_precomputed=True)
        assert np.array_equal(out2, _expect(oracle_mod, c3, 1 << 16))


@pytest.mark.parametrize("n", [1_000_000, 16_000_000])
def test_full_size_properties(gs, oracle_mod, n):
    """BASELINE sizes: permutation, monotone buckets, reverse-stable ties; 1M is also diffed against the oracle."""
    c = cases.sort_case(seed=10, n=n, index_kind="shuffled" if n <= 1_000_000 else "identity")
    R = 1 << 16
    with gs.Engine(n) as e:
        e.upload_centers(c["centers"])
        out, ms = e.sort(c["mvp"], n, n, c["indexes"])
    assert np.array_equal(np.sort(out), np.arange(n, dtype=np.uint32)), "not a permutation"
    # recompute buckets independently (numpy, int64 wrap) and check the ordering contract
    m = c["mvp"]
    rows = np.array([int(np.float64(m[2]) * 1000.0), int(np.float64(m[6]) * 1000.0), int(np.float64(m[10]) * 1000.0)], np.int64)
    d = (c["centers"][:, :3].astype(np.int64) * rows).sum(1)
    d = ((d + 2**31) % 2**32 - 2**31).astype(np.int32)
    dmin, dmax = int(d.min()), int(d.max())
    rm = np.float32(R - 1) / (np.float32(dmax) - np.float32(dmin))
    b = ((d.astype(np.int64) - dmin).astype(np.float32) * rm).astype(np.int32)
    bs = b[out]
    assert np.all(np.diff(bs) <= 0), "buckets must be non-increasing (back to front)"
    pos = np.empty(n, np.int64)
    pos[c["indexes"]] = np.arange(n)
    same = np.diff(bs) == 0
    assert np.all(np.diff(pos[out])[same] < 0), "equal buckets must come in descending input position"
    if n <= 1_000_000:
        assert np.array_equal(out, _expect(oracle_mod, c, R))


def test_edge_cases(gs, oracle_mod):
    # single splat and all-equal distances: the reference traps (NaN bucket); we define bucket 0 -> reversed input order
    c = cases.sort_case(seed=1, n=1)
    assert np.array_equal(gs.sort_indexes(*cases.call_args(c, 1 << 16)), c["indexes"])
    n = 5000
    c = cases.sort_case(seed=2, n=n)
    c["centers"][:] = c["centers"][0]
    got = gs.sort_indexes(*cases.call_args(c, 1 << 16))
    assert np.array_equal(got, c["indexes"][::-1])
    # sortCount = 0: pure copy-through
    c = cases.sort_case(seed=3, n=1000, sort_frac=0.0)
    assert np.array_equal(gs.sort_indexes(*cases.call_args(c, 1 << 16)), c["indexes"])
    # bad arguments are reported, not executed
    with pytest.raises(gs.GsError):
        gs.sort_indexes(c["indexes"], c["centers"], None, c["mvp"], None, None, 1 << 16, 2000, 1000, 1000, False, True, False)


def test_float_24bit_precision_overshoot_is_clamped(gs, oracle_mod):
    """Float mode allows 24-bit precision (Viewer.js:208-210); there (R-1)/range * (max-min) can round to exactly R for the
    farthest splat.  The reference then increments frequencies[R] (outside its prefix sum); we clamp to R-1."""
    R = 1 << 24
    hit = 0
    for seed in range(6):
        c = cases.sort_case(seed=seed, n=20000, integer=False)
        got = gs.sort_indexes(*cases.call_args(c, R))
        want, buckets = oracle_mod.port_sort_indexes(*cases.call_args(c, R), want_buckets=True)
        assert np.array_equal(got, want)
        assert np.array_equal(np.sort(got), np.sort(c["indexes"]))
        hit += int(buckets.max() == R - 1)
    assert hit > 0


def test_engine_argument_errors_and_empty_work(gs):
    """Error behaviour of the boundary: bad arguments are reported with a status, nothing is executed, nothing crashes."""
    c = cases.sort_case(seed=4, n=1000)
    with gs.Engine(1000) as e:
        with pytest.raises(gs.GsError):                       # more centres than the engine was created for
            e.upload_centers(np.zeros((2000, 4), np.int32))
        e.upload_centers(c["centers"][:400])
        out, _ = e.sort(c["mvp"], 1000, 1000, None)           # counts clamp to what has been uploaded (SortWorker.js:99-100)
        assert out.shape[0] == 1000 and np.array_equal(np.sort(out[:400]), np.arange(400, dtype=np.uint32))
        out0, _ = e.sort(c["mvp"], 0, 0, None)                 # nothing to sort
        assert out0.shape[0] == 0
        with pytest.raises(gs.GsError):                       # no framebuffer was requested at create time
            from gaussiansplats3d_b200.engine import Uniforms
            e.render(Uniforms(np.eye(4), np.eye(4), np.zeros(3), (1, 1), (8, 8)), 8, 8, 10)
    with pytest.raises(gs.GsError):
        gs.Engine(10, distance_map_range=1)                   # range < 2


@pytest.mark.parametrize("integer,dynamic", [(True, False), (True, True), (False, False), (False, True)])
def test_distance_prepass_all_four_shader_variants(gs, oracle_mod, integer, dynamic):
    """D1: SplatMesh.computeDistancesOnGPU, the four transform-feedback vertex shaders (SplatMesh.js:1451-1502) with their host side
    (:1722-1738, getIntegerMatrixArray :2057-2064 = Math.round(element * 1000)).  Integer variants are bit-exact (wrapping int32);
    the float variants are GLSL f32 dot products whose summation order a driver may choose, so they are compared to 1 ulp of the
    largest term.  The distances then drive the sorter's precomputed branch (sorter.cpp:31-38, 79-86) bit-exactly."""
    n = 50_000
    c = cases.sort_case(seed=6, n=n, integer=integer, dynamic=dynamic, index_kind="identity")
    mvp64 = c["mvp"].astype(np.float64)
    T64 = None
    if dynamic:
        T64 = np.zeros((32, 16), np.float64)
        T64[:] = np.eye(4).T.reshape(16)
        T64[: c["transforms"].shape[0]] = c["transforms"].astype(np.float64)
    with gs.Engine(n, integer_based_sort=integer, dynamic_mode=dynamic) as e:
        e.upload_centers(c["centers"], c["scene_indexes"])
        d = e.compute_distances(mvp64, n, T64)
        M = mvp64.reshape(4, 4).T
        if dynamic:     # tempMatrix = mvp * transform_s (premultiply), row 2 of it: elements [2], [6], [10], [14]
            rows = np.stack([(M @ T64[s].reshape(4, 4).T)[2] for s in range(32)])
        else:
            rows = np.tile(M[2], (32, 1))
        scene = c["scene_indexes"].astype(np.int64) if dynamic else np.zeros(n, np.int64)
        if integer:
            ir = np.floor(rows * 1000.0 + 0.5).astype(np.int64)[scene]
            cc = c["centers"].astype(np.int64)
            want = cc[:, 0] * ir[:, 0] + cc[:, 1] * ir[:, 1] + cc[:, 2] * ir[:, 2] + (ir[:, 3] * cc[:, 3] if dynamic else 0)
            want = ((want + 2**31) % 2**32 - 2**31).astype(np.int32)
            assert d.dtype == np.int32 and np.array_equal(d, want)
        else:
            fr = rows.astype(np.float32)[scene]
            cc = c["centers"].astype(np.float32)
            terms = np.stack([cc[:, 0] * fr[:, 0], cc[:, 1] * fr[:, 1], cc[:, 2] * fr[:, 2]] + ([fr[:, 3]] if dynamic else []), 1)
            want = terms.astype(np.float64).sum(1)
            assert d.dtype == np.float32
            # every partial sum is rounded once: a few ulps of the magnitude being summed
            assert np.all(np.abs(d.astype(np.float64) - want) <= 4.0 * np.spacing(np.abs(terms).sum(1).astype(np.float32)).astype(np.float64) + 1e-30)
        out, _ = e.sort(c["mvp"], n, n, None, precomputed=d, transforms=c["transforms"])
        c3 = dict(c, precomputed=d, use_precomputed=True)
        assert np.array_equal(out, _expect(oracle_mod, c3, 1 << 16))


def test_void_twin_sortIndexes_has_the_reference_signature(gs, oracle_mod):
    """`extern "C" void sortIndexes(...16 args...)` (sorter.cpp:17-22): the symbol a wasm-free host would bind in place of the reference's;
    same arguments, no return value, indexesOut untouched on failure (like a trap that aborts the call)."""
    import ctypes as C
    lib = gs._native.load()
    c = cases.sort_case(seed=3, n=30_000, index_kind="octree", sort_frac=0.6)
    R = 1 << 16
    out = np.full(c["render_count"], 0xFFFFFFFF, np.uint32)
    mapped = np.zeros(c["render_count"], np.int32)
    freq = np.zeros(2 * R, np.uint32)
    idx = np.ascontiguousarray(c["indexes"], np.uint32); cen = np.ascontiguousarray(c["centers"]); mvp = np.ascontiguousarray(c["mvp"], np.float32)
    lib.sortIndexes.restype = None
    lib.sortIndexes.argtypes = [C.c_void_p] * 9 + [C.c_uint32] * 4 + [C.c_bool] * 3
    p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)   # noqa: E731
    lib.sortIndexes(p(idx), p(cen), None, p(mapped), p(freq), p(mvp), p(out), None, None, R, c["sort_count"], c["render_count"], c["splat_count"], False, True, False)
    assert np.array_equal(out, _expect(oracle_mod, c, R))
    # failure path: sortCount > renderCount is refused, the output buffer keeps its contents
    sentinel = np.full(c["render_count"], 123456789, np.uint32)
    lib.sortIndexes(p(idx), p(cen), None, None, None, p(mvp), p(sentinel), None, None, R, c["render_count"] + 1, c["render_count"], c["splat_count"], False, True, False)
    assert (sentinel == 123456789).all()
    lib.gs_dropin_release()          # the cached private engine of the stateless entry can be dropped explicitly
    lib.sortIndexes(p(idx), p(cen), None, p(mapped), p(freq), p(mvp), p(out), None, None, R, c["sort_count"], c["render_count"], c["splat_count"], False, True, False)
    assert np.array_equal(out, _expect(oracle_mod, c, R))
