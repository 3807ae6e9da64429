"""CPU: the C-ABI library builds, loads and exports every symbol include/gsplat_b200.h declares; without a GPU every
compute entry fails loudly (there is no CPU fallback in the product)."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    from gaussiansplats3d_b200 import build, _native
    build.build()
    return _native.load()


def declared_symbols():
    text = (ROOT / "include" / "gsplat_b200.h").read_text()
    return sorted(set(re.findall(r"GS_API\s+[\w\s\*]+?\b(\w+)\s*\(", text)))


def test_header_declares_expected_entry_points():
    syms = declared_symbols()
    for s in ("gs_sort_indexes", "sortIndexes", "gs_create", "gs_sort", "gs_render", "gs_frame", "gs_upload_centers", "gs_upload_splat_data"):
        assert s in syms


@pytest.mark.parametrize("sym", declared_symbols())
def test_library_exports_symbol(lib, sym):
    assert hasattr(lib, sym), f"{sym} declared in include/gsplat_b200.h but not exported"


def test_python_binding_lists_every_symbol():
    from gaussiansplats3d_b200 import _native
    assert sorted(_native.EXPORTED_SYMBOLS) == declared_symbols()


def test_struct_sizes_match_header(lib):
    """ctypes mirrors must match the C layout: compile a tiny C probe with the real header."""
    import subprocess, tempfile
    from gaussiansplats3d_b200 import _native as N
    src = '#include "gsplat_b200.h"\n#include <stdio.h>\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu\\n",sizeof(gs_config),sizeof(gs_sort_params),sizeof(gs_splat_data),sizeof(gs_uniforms),sizeof(gs_render_params),sizeof(gs_projected_splat),sizeof(gs_timings));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        p = Path(d) / "probe.c"
        p.write_text(src)
        subprocess.run(["/usr/bin/gcc", "-I", str(ROOT / "include"), str(p), "-o", str(Path(d) / "probe")], check=True)
        out = subprocess.run([str(Path(d) / "probe")], capture_output=True, text=True, check=True).stdout.split()
    want = [C.sizeof(t) for t in (N.gs_config, N.gs_sort_params, N.gs_splat_data, N.gs_uniforms, N.gs_render_params, N.gs_projected_splat, N.gs_timings)]
    assert [int(v) for v in out] == want


def test_no_cpu_fallback_without_device(lib):
    from gaussiansplats3d_b200 import Engine, GsError
    if lib.gs_device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(GsError) as ei:
        Engine(1000)
    assert ei.value.code == 2  # GS_ERR_NO_DEVICE
    from gaussiansplats3d_b200 import sort_indexes
    with pytest.raises(GsError):
        sort_indexes(np.arange(4, dtype=np.uint32), np.zeros((4, 4), np.int32), None, np.eye(4, dtype=np.float32).reshape(16), None, None,
                     1 << 16, 4, 4, 4, False, True, False)


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under gaussiansplats3d_b200/ may reference it."""
    for p in (ROOT / "gaussiansplats3d_b200").rglob("*"):
        if p.suffix in (".py", ".cu", ".cuh", ".h") and p.is_file():
            t = p.read_text()
            assert "import oracle" not in t and "from oracle" not in t and "oracle/" not in t.replace("oracle/_ref", ""), p


def test_napi_addon_type_checks_and_binds_every_entry_point():
    """js/gsplat_b200_addon.cc (the Node.js side of the drop-in; Node itself is absent from this image): compiles cleanly against the
    declarations of the Node-API it uses (js/test/node_api_stub.h) with -Wall -Wextra, calls every GS_API symbol of the header, and
    exports one JS function per symbol; the JS shims only call functions the addon exports."""
    import subprocess
    res = subprocess.run(["/usr/bin/g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Wextra", "-Werror", "-DGS_NAPI_STUB", str(ROOT / "js" / "gsplat_b200_addon.cc")],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    src = (ROOT / "js" / "gsplat_b200_addon.cc").read_text()
    for sym in declared_symbols():
        assert re.search(r"\b" + sym + r"\s*\(", src), f"{sym} is not called by the addon"
    exported = set(re.findall(r'EXPORT\("(\w+)"', src))
    assert len(exported) >= len(declared_symbols())
    for shim in ("SortWorkerB200.js", "SplatMeshB200.js"):
        used = set(re.findall(r"\baddon\.(\w+)\(", (ROOT / "js" / shim).read_text()))
        assert used and used <= exported, (shim, used - exported)
    # the worker shim speaks the whole protocol of src/worker/SortWorker.js
    w = (ROOT / "js" / "SortWorkerB200.js").read_text()
    for key in ("sortSetupPhase1Complete", "indexesToSortBuffer", "sortedIndexesBuffer", "precomputedDistancesBuffer", "transformsBuffer", "sortDone",
                "sortCanceled", "splatSortCount", "splatRenderCount", "usePrecomputedDistances", "precomputedDistances", "transforms", "sceneIndexes"):
        assert key in w, key
